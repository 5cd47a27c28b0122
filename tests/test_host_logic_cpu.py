"""CPU tests of round-2 host logic: the segmented-Jacobian decision, the by-value segment table's layout against the
header, the pre-replay callback registry, and bench.py's arm-independent workload string."""
import ctypes
import os
import re
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_segment_table_layout_matches_the_header():
    from movae_b200 import _lib as L

    hdr = open(os.path.join(ROOT, "include", "movae_b200.h")).read()
    assert int(re.search(r"#define MOVAE_MAX_SEGMENTS (\d+)", hdr).group(1)) == L.MAX_SEGMENTS
    assert int(re.search(r"#define MOVAE_MAX_K (\d+)", hdr).group(1)) == L.MAX_K
    # int32 n_segments, int32 k, int64 n[S], int64 out_off[S], pointer rows[S][MAX_K]: passed to the kernel BY VALUE (< 4 KB)
    assert ctypes.sizeof(L.JacSegments) == 8 + 8 * L.MAX_SEGMENTS * 2 + 8 * L.MAX_SEGMENTS * L.MAX_K
    assert ctypes.sizeof(L.JacSegments) < 4096 - 256
    assert L.JacSegments.rows.offset == 8 + 16 * L.MAX_SEGMENTS
    assert int(re.search(r"#define MOVAE_ABI_VERSION (\d+)", hdr).group(1)) == L.ABI_VERSION


def test_segmented_path_is_chosen_only_when_nobody_needs_a_matrix():
    import movae_b200 as mv
    from movae_b200 import _lib as L
    from movae_b200 import autojac

    params = [torch.nn.Parameter(torch.zeros(3)) for _ in range(4)]
    agg = mv.UPGrad()
    assert autojac._segments_possible(params, agg)
    many = [torch.nn.Parameter(torch.zeros(1)) for _ in range(L.MAX_SEGMENTS + 1)]
    assert not autojac._segments_possible(many, agg)                       # more tensors than the by-value table holds
    h = agg.weighting.register_forward_hook(lambda m, i, o: None)          # hooks receive J (main.py:1248-1250)
    assert not autojac._segments_possible(params, agg)
    h.remove()
    assert autojac._segments_possible(params, agg)
    agg.weighting.gramian_reducer = lambda G: None                         # a torch.distributed reducer sits between K1 and K2
    assert not autojac._segments_possible(params, agg)
    agg.weighting.gramian_reducer = None
    agg.data_parallel = object()                                           # rows are reduce-scattered: needs the flat J
    assert not autojac._segments_possible(params, agg)
    comfort = mv.COMFORT()
    assert autojac._segments_possible(params, comfort) and comfort.supports_segments()


def test_pre_replay_callbacks_need_a_graphed_step():
    from movae_b200 import optim

    with pytest.raises(RuntimeError, match="GraphedStep"):
        optim.register_pre_replay(lambda: None)


def test_bench_workload_string_is_arm_independent_and_names_the_config():
    sys.path.insert(0, ROOT)
    import bench

    w = bench.workload_name(3, 100_000_000, "upgrad", "weak")
    assert w == "aggregation microbench k=3 P=100000000 per GPU agg=upgrad (BASELINE.json configs[4])"
    assert "global" in bench.workload_name(3, 100_000_000, "upgrad", "strong")
    assert bench.algorithmic_bytes(3, 100_000_000) == {"gram": 1_200_000_000, "recombine": 1_600_000_000, "step": 2_800_000_000}
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("workload_name(") >= 3                                 # both arms build config.workload through it


def test_comfort_and_pnupgrad_report_their_device_side_state_in_the_spec():
    import movae_b200 as mv
    from movae_b200 import _lib as L

    c = mv.COMFORT(mgda_norm_type="l2")
    spec, vec, aux = c.weighting.solve_spec(3)
    assert spec.kind == L.SOLVE_MGDA and aux is None                        # no blend coefficients yet: plain MGDA semantics
    p = mv.PNUPGrad(prob=0.3)
    spec, vec, aux = p.weighting.solve_spec(3)
    assert spec.kind == L.SOLVE_UPGRAD and spec.mode == L.UPGRAD_NORM["draw"] and aux is None   # the flag exists after the first CUDA call
    torch.manual_seed(5)
    draws = [torch.rand(1).item() < 0.3 for _ in range(5)]
    torch.manual_seed(5)
    assert [p.weighting.draw() == "l2" for _ in range(5)] == draws          # host RNG consumption as pnupgrad.py:129
