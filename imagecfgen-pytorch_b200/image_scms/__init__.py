"""Drop-in mirror of the reference's ``image_scms`` package (same module, class and function names) whose
models execute in the B200-native CUDA extension.  Import it with ``imagecfgen-pytorch_b200`` on sys.path:

    import sys; sys.path.insert(0, "imagecfgen-pytorch_b200")
    from image_scms import mnist
"""
