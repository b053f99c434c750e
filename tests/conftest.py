import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The CPU suite needs the two native libraries (C-ABI .so to check its exports, oracle .so as the checker).
    Both cross-compile without a GPU; build them if a fresh checkout has not run __graft_entry__.build() yet."""
    import subprocess
    lib = os.path.join(ROOT, "luciddreamer_b200", "csrc", "libgsraster_b200.so")
    if not os.path.exists(lib):
        subprocess.check_call(["make", "-C", os.path.dirname(lib), "-j8"], stdout=subprocess.DEVNULL)
    bind = os.path.join(os.path.dirname(lib), "_gsraster_torch.so")
    if not os.path.exists(bind):                     # host-side torch binding over the same C ABI (plain g++, ~30 s)
        import sys
        subprocess.check_call([sys.executable, os.path.join(os.path.dirname(lib), "build_binding.py")],
                              stdout=subprocess.DEVNULL)
    ora = os.path.join(ROOT, "oracle", "liboracle.so")
    if not os.path.exists(ora):
        subprocess.check_call(["make", "-C", os.path.dirname(ora), "liboracle.so"], stdout=subprocess.DEVNULL)
