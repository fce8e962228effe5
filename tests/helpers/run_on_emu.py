"""TEST HARNESS (CPU): run one of the repo's entry points on the host emulation (emu_device.py).

    python tests/helpers/run_on_emu.py smoke
    python tests/helpers/run_on_emu.py bench [--particles K] [bench.py arguments]
    python tests/helpers/run_on_emu.py script path/to/script.py [arguments]

A dry run of the Python around the kernels: numbers printed by it are meaningless."""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)


def main():
    import emu_device
    emu_device.install()
    what, rest = sys.argv[1], sys.argv[2:]
    if what == "smoke":
        import __graft_entry__
        __graft_entry__.smoke()
    elif what == "bench":
        from mjmpc_b200 import _lib

        def fake_peak(device, blocks, iters, tflops, ms):       # the probe measures hardware
            tflops._obj.value, ms._obj.value = 1.0, 1.0
            return 0
        _lib.lib().mjb_fp64_peak = fake_peak
        import bench
        if "--particles" in rest:
            i = rest.index("--particles")
            bench.K_GLOBAL = int(rest[i + 1])
            del rest[i:i + 2]
        sys.argv = ["bench.py"] + rest
        bench.main()
    elif what == "script":
        sys.argv = rest
        runpy.run_path(rest[0], run_name="__main__")
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
