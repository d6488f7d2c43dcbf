"""CPU-only checks of bench.py's work model: the algorithmic FLOPs / bytes that `roofline.achieved` is computed from are
SURVEY.md 8d's per-config figures, and the strong-scaling split is the one BASELINE configs[4] names."""
import pytest

import bench


# SURVEY.md 8d "Per-config totals": embed GFLOP per cloud as written / minimal, compulsory HBM bytes per batch
SURVEY_8D = {
    "c1": (3.908, 2.331, 2.5e6),
    "c2": (9.945, 7.604, 28e6),
    "c3": (31.26, 18.65, 94e6),
    "c4": (318.2, 242.1, 63e6),
}


@pytest.mark.parametrize("name", list(SURVEY_8D))
def test_algorithmic_work_is_the_surveys(name):
    w = bench.WORKLOADS[name]
    a = bench.algorithmic_work(w)
    as_written, minimal, batch_bytes = SURVEY_8D[name]
    assert a["embed_flops_as_written"] / 1e9 == pytest.approx(as_written, rel=2e-3)
    assert a["embed_flops"] / 1e9 == pytest.approx(minimal, rel=2e-3)
    if name == "c3":        # the survey counts C3's tokens as bf16 (2 B); bench.py counts the fp32 tokens the module returns
        batch_bytes += 2 * 512 * 256 * w["B"]
    assert a["compulsory_bytes"] * w["B"] == pytest.approx(batch_bytes, rel=0.05)
    # the executed count may exceed the minimal one only by the padded reduction widths - never reach the as-written one
    assert a["embed_flops"] * 0.99 <= a["embed_flops_executed"] < a["embed_flops_as_written"]
    assert a["fps_pairs"] == a["knn_pairs"] > 0


def test_c2_batch_totals_match_the_verdicts_recomputation():
    """Round-1 verdict: 524 288 points x 919 040 MAC + 16 384 groups x 294 912 MAC = 0.9733 TFLOP per 128 clouds."""
    w = bench.WORKLOADS["c2"]
    total = bench.algorithmic_work(w)["embed_flops"] * w["B"]
    assert total == 2 * (524288 * 919040 + 16384 * 294912)
    assert total / 1e12 == pytest.approx(0.9733, rel=1e-3)


def test_strong_scaling_split_and_weak_default():
    c5 = bench.WORKLOADS["c5"]
    assert [bench.per_gpu_clouds(c5, n) for n in (1, 2, 4, 8)] == [4096, 2048, 1024, 512]     # BASELINE configs[4]: B/P per GPU
    with pytest.raises(SystemExit):
        bench.per_gpu_clouds(c5, 3)
    assert all(bench.per_gpu_clouds(bench.WORKLOADS["c2"], n) == 128 for n in (1, 2, 8))      # weak: fixed per-GPU batch


def test_peaks_come_from_the_measured_file_or_the_stated_fallback():
    p = bench.peaks()
    assert p["source"].startswith(("measured", "fallback"))
    assert 3000 < p["hbm_gbs"] < 9000 and 800 < p["bf16_tflops_sustained"] <= p["bf16_tflops"] < 2600


def test_traffic_is_measured_or_null_with_a_reason():
    v, why = bench.measured_traffic("c2", "bf16", 128)
    assert (v is None and why) or (v > 28e6 and why)          # never below the compulsory bytes
    v, why = bench.measured_traffic("no-such-workload", "bf16", 1)
    assert v is None and "no ncu capture" in why
