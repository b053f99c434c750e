"""Builds the torch binding (_gsraster_torch.so) in-tree with plain g++: torch_binding.cpp is host-only C++ that links
against libgsraster_b200.so (the CUDA product) and the torch libraries of the running interpreter.
Usage: python build_binding.py   (called by the Makefile / __graft_entry__.build())"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))


def main() -> int:
    import torch
    from torch.utils import cpp_extension as ce
    src = os.path.join(HERE, "torch_binding.cpp")
    out = os.path.join(HERE, "_gsraster_torch.so")
    deps = [src, os.path.join(HERE, "..", "..", "include", "gsraster.h"), os.path.join(HERE, "libgsraster_b200.so")]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps if os.path.exists(d)):
        return 0
    cuda_home = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    inc = ce.include_paths() + [sysconfig.get_paths()["include"], os.path.join(HERE, "..", "..", "include"),
                                os.path.join(cuda_home, "include")]
    libdir = ce.library_paths()[0]
    cxx = os.environ.get("GS_CXX") or "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-w", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           "-DTORCH_EXTENSION_NAME=_gsraster_torch", "-DTORCH_API_INCLUDE_EXTENSION_H"]
    cmd += [f"-I{p}" for p in inc]
    cmd += [src, "-o", out, f"-L{libdir}", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", "-ltorch_python",
            f"-L{HERE}", "-lgsraster_b200", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{libdir}"]
    print(" ".join(cmd))
    return subprocess.call(cmd)


if __name__ == "__main__":
    sys.exit(main())
