"""oracle/build_ref.py -- TEST INFRASTRUCTURE ONLY; build container only.

Compiles the REFERENCE's own hash-grid CUDA extension, unmodified, from the sources where they
lie under /root/reference (encoder/hashencoder/src/{hashencoder.cu,bindings.cpp}) into
oracle/_ref/ (git-ignored; travels to the GPU box with the snapshot).  The only change is the
language flag: the reference asks for -std=c++14 (encoder/hashencoder/backend.py:7-12), which
torch >= 2.1 headers reject, so -std=c++17 is used.  Nothing is copied into the repo.

The resulting module `_ref_hash_encoder` is the GPU-side pin of the hash encoder: tests/
test_gpu_reference_kernel.py runs it next to libavatarcraft_b200.so on the same inputs.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/encoder/hashencoder/src"
SRC_RM = "/root/reference/raymarching/src"
OUT = os.path.join(ROOT, "oracle", "_ref")


def main():
    if not os.path.isdir(SRC):
        print("build_ref: /root/reference not present, skipping")
        return
    target = os.path.join(OUT, "_ref_hash_encoder.so")
    if os.path.exists(target):
        print("build_ref: up to date")
        return
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    load(name="_ref_hash_encoder", sources=[os.path.join(SRC, "hashencoder.cu"), os.path.join(SRC, "bindings.cpp")],
         extra_cflags=["-O3", "-std=c++17"],
         extra_cuda_cflags=["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                            "-U__CUDA_NO_HALF2_OPERATORS__"],
         build_directory=OUT, verbose=False, is_python_module=False)
    print("build_ref: built", target)


def build_raymarching():
    """The reference's raymarching extension (dead code there, but its kernels are the oracle of ours)."""
    target = os.path.join(OUT, "_ref_raymarching.so")
    if not os.path.isdir(SRC_RM) or os.path.exists(target):
        return
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    load(name="_ref_raymarching", sources=[os.path.join(SRC_RM, "raymarching.cu"), os.path.join(SRC_RM, "bindings.cpp")],
         extra_cflags=["-O3", "-std=c++17"], extra_cuda_cflags=["-O3", "-std=c++17"], build_directory=OUT, verbose=False,
         is_python_module=False)
    for f in os.listdir(OUT):
        if f.endswith((".o", ".d")) or f in ("lock", "build.ninja", ".ninja_deps", ".ninja_log"):
            os.remove(os.path.join(OUT, f))
    print("build_ref: built", target)


def build_shencoder():
    """The reference's spherical-harmonics extension (encoder/shencoder/src/{shencoder.cu,bindings.cpp}): the GPU pin of
    ac_sh_encode_forward / _backward (tests/test_gpu_encoders.py)."""
    src = "/root/reference/encoder/shencoder/src"
    target = os.path.join(OUT, "_ref_sh_encoder.so")
    if not os.path.isdir(src) or os.path.exists(target):
        return
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    load(name="_ref_sh_encoder", sources=[os.path.join(src, "shencoder.cu"), os.path.join(src, "bindings.cpp")],
         extra_cflags=["-O3", "-std=c++17"],
         extra_cuda_cflags=["-O3", "-std=c++17", "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                            "-U__CUDA_NO_HALF2_OPERATORS__"],
         build_directory=OUT, verbose=False, is_python_module=False)
    for f in os.listdir(OUT):
        if f.endswith((".o", ".d")) or f in ("lock", "build.ninja", ".ninja_deps", ".ninja_log"):
            os.remove(os.path.join(OUT, f))
    print("build_ref: built", target)


if __name__ == "__main__":
    main()
    build_raymarching()
    build_shencoder()
