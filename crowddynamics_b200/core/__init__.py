"""Functional API with the reference's module paths and function names (crowddynamics/core/...): every function
takes the structured ``agents`` array, mutates it in place and runs on the GPU (strict mode: upload, kernel, download)."""
