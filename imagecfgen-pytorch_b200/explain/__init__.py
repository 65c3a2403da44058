"""Drop-in import path of the reference's ``explain`` package for the two generator-driven explainers."""
