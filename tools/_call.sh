# scratch command file of the build sessions: `gpurun -- 'bash tools/_call.sh'` runs whatever the session last wrote here
python -m pytest tests -m gpu -q -x 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --skip-cpu --skip-torch > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
