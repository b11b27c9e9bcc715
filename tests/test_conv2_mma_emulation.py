"""CPU: index logic of the mma.sync conv2 kernels (conv2_mma.cu) -- per-lane fragment assembly emulated in Python around
an m16n8k8 MMA with the documented fragment layouts, against torch conv3d / autograd (scripts/emulate_conv2_mma.py).
The GPU parity tests (test_policy_gpu.py, GNBV_CONV2_TC=2/6) cover the kernels themselves."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fragment_mapping_reproduces_conv3d_and_its_data_gradient():
    spec = importlib.util.spec_from_file_location("emulate_conv2_mma", os.path.join(ROOT, "scripts", "emulate_conv2_mma.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m.check(9, B=2)        # reference-native 20^3 grid: G1 = 9, G2 = 4 (partial tiles, every parity class)
    m.check(10, B=1)       # even G1: equal-sized parity classes, out-of-range odd taps
    m.check_conv1(20, B=1)  # conv1 forward: slab offsets of the 32 k slots, two-row fragments, partial z tiles
    m.check_conv1_wgrad(20, B=1)   # conv1 weight gradient: position k-steps, tap columns, per-warp partial records
