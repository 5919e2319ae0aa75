"""bench.py's per-stage roofline arithmetic on the committed live stage table (CPU only: no GPU, no product library)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _stage_table():
    rows = []
    for line in open(os.path.join(ROOT, "profiles", "r02_stage_times_b256.txt")):
        p = line.split()
        if len(p) == 2 and p[0] not in ("TOTAL", "ok"):
            try:
                rows.append((p[0], float(p[1])))
            except ValueError:
                pass
    return rows


def test_stage_roofline_of_the_committed_stage_table():
    b = _bench()
    from birdnet_b200.modelgen import get_spec
    spec = get_spec("birdnet_v24")
    st = _stage_table()
    assert len(st) >= 30 and any(n == "spectrogram" for n, _ in st)
    peaks = {"hbm": 6546.2, "tf_sust": 1409.0, "tf_burst": 1500.0, "src": "test"}
    dom, stages = b._stage_roofline(spec, st, 256, peaks)
    by = {s["stage"]: s for s in stages}
    # the fused front-end kernel: tensor-bound on ALGORITHMIC flops, the three-product policy reported beside it
    fe = by["spectrogram"]
    assert fe["bound"] == "tensor" and fe["unit"] == "TFLOP/s" and fe["products_per_mac"] == 3
    assert abs(fe["frac_counting_3_products"] - 3 * fe["frac"]) < 1e-9
    assert 0.05 < fe["frac"] < 1.0 and fe["ceilings_ms"]["tensor_3_products"] > fe["ceilings_ms"]["hbm"]
    # 151 M MAC per segment: 2 flops each
    assert abs(fe["alg_per_segment"] / 2 - 150.7e6) < 1e6
    # every fused MBConv block carries its ceilings and names the slowest one as its bound
    mb = [s for s in stages if s["stage"].endswith(".mbconv")]
    assert len(mb) == 5 and all(s["bound"] == max(s["ceilings_ms"], key=s["ceilings_ms"].get) for s in mb)
    # fractions are achieved / peak everywhere and no stage claims more than its peak
    for s in stages:
        assert abs(s["frac"] - s["achieved"] / s["peak"]) < 1e-9 and 0 < s["frac"] < 1.0, s
    assert dom["stage"] == max(stages, key=lambda x: x["ms"])["stage"]


def test_kernel_classes_cover_every_stage_name():
    b = _bench()
    names = {b._kernel_class(n) for n, _ in _stage_table()}
    assert any("k_spec_v24" in c for c in names) and any("k_mbconv" in c for c in names)
    assert any("k_dw_se" in c for c in names) and any("k_stem_planes" in c for c in names)
