"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol the
header declares, the ctypes mirror matches the C struct, and the product fails loudly (no
fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "dagma_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dagma_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    if not os.path.exists(__graft_entry__.LIB):
        __graft_entry__.build()
    from midagma_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dagma_b200.h but not exported"
    for n in _lib.EXPORTS:
        assert n in names, f"{n} bound in _lib.py but not declared in the header"
    assert lib.dagma_version() >= 100
    assert lib.dagma_last_error() is not None


def test_struct_mirror_layout():
    from midagma_b200 import _lib
    a = _lib.SmallFitArgs
    # 6 int32 (24 B), 4 doubles, 2 x 16 doubles, 16 int32, 12 pointers
    assert C.sizeof(a) == 24 + 4 * 8 + 2 * 16 * 8 + 16 * 4 + 12 * 8
    assert a.mu.offset == 56 and a.iters.offset == 56 + 256 and a.cov.offset == 56 + 256 + 64


@pytest.mark.skipif(torch.cuda.is_available(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    from midagma_b200 import DagmaLinear, fit_batch, _lib
    from midagma_b200.nonlinear import DagmaMLP
    X = np.random.default_rng(0).normal(size=(50, 5))
    with pytest.raises(_lib.DagmaB200Error):
        DagmaLinear("l2").fit(X)
    with pytest.raises(_lib.DagmaB200Error):
        fit_batch(X[None])
    with pytest.raises((_lib.DagmaB200Error, AssertionError, RuntimeError)):
        DagmaMLP([5, 3, 1]).h_func()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "midagma_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def test_structured_logger_sinks(tmp_path):
    """midagma_b200.logger mirrors the fork's src/logger.py interface: rows in memory, jsonl / csv files, callback."""
    import json
    from midagma_b200.logger import LogConfig, StructuredLogger, build_default_logger
    seen = []
    cfg = LogConfig(enabled=True, store_jsonl=True, store_csv=True, run_dir=str(tmp_path / "run"), run_name="t",
                    meta={"k": 1}, callback=seen.append)
    lg = StructuredLogger(build_default_logger(name="t_logger", stream=False), cfg)
    lg.emit("minimize.checkpoint", {"iter": 100, "obj_total": 1.5, "reg_dag_cfg": {"s": 1.0}})
    lg.emit("other", {"iter": 200, "obj_total": 2.5, "reg_dag_cfg": {"s": 0.9}})
    cols = lg.load(event="minimize.checkpoint")
    assert list(cols["iter"]) == [100] and seen[1]["event"] == "other"
    lg.close()
    lines = [json.loads(x) for x in open(tmp_path / "run" / "metrics.jsonl")]
    assert [r["iter"] for r in lines] == [100, 200]
    assert json.load(open(tmp_path / "run" / "meta.json"))["k"] == 1
    assert open(tmp_path / "run" / "metrics.csv").read().splitlines()[0].startswith("event,iter,obj_total")
    off = StructuredLogger(build_default_logger(name="t_logger", stream=False), LogConfig(enabled=False))
    off.emit("x", {"a": 1})
    assert off._rows == [] and off.run_dir is None


def test_mlp_without_bias_fails_like_the_reference():
    """The reference's DagmaMLP cannot be built with bias=False: its constructor calls nn.init.zeros_(self.fc1.bias)
    unconditionally (src/dagma/nonlinear.py:36-38) and dies with AttributeError on None.  The drop-in keeps that --
    it must not quietly accept a configuration the reference rejects.  LocallyConnected alone does take bias=False
    (locally_connected.py:38-43; GPU test in tests/test_scale_gpu.py)."""
    from midagma_b200.nonlinear import DagmaMLP, LocallyConnected
    with pytest.raises(AttributeError):
        DagmaMLP([5, 4, 1], bias=False)
    lc = LocallyConnected(5, 4, 3, bias=False)
    assert lc.bias is None and lc.weight.shape == (5, 4, 3)


def test_batch_lanes_and_value_checks_on_cpu(monkeypatch):
    """Host-side decisions that need no device: how many problems of a mid-d batch run side by side, and the input
    checks of utils.count_accuracy (src/dagma/utils.py:269-278)."""
    import numpy as np
    from midagma_b200 import utils
    from midagma_b200.linear import _batch_lanes
    monkeypatch.delenv("DAGMA_BATCH_LANES", raising=False)
    assert _batch_lanes(200, 16) == 1                      # beyond d = 128 the blocked inverse fills the device
    monkeypatch.setenv("DAGMA_BATCH_LANES", "3")
    assert _batch_lanes(100, 16) == 3 and _batch_lanes(100, 2) == 2
    B = np.zeros((4, 4), dtype=np.int64)
    B[0, 1] = 1
    assert utils._check_values(B) is False
    B[1, 2] = -1
    assert utils._check_values(B) is True
    B[2, 1] = -1
    with pytest.raises(ValueError, match="only appear once"):
        utils._check_values(B)
    with pytest.raises(ValueError, match="value in"):
        utils._check_values(2 * np.abs(B))


def test_pipeline_chunks():
    """Chunks of the copy / compute pipeline of minimize_batch: whole waves, contiguous, complete, no sliver."""
    from midagma_b200.linear import _pipeline_chunks
    w = 296
    assert _pipeline_chunks(4096, w) == [(0, 592), (592, 1776), (1776, 2960), (2960, 4096)]
    for batch in (4 * w, 4 * w + 1, 5 * w - 1, 6 * w, 6 * w + 7, 10 * w + 295, 32768, 1):
        b = _pipeline_chunks(batch, w)
        assert b[0][0] == 0 and b[-1][1] == batch and all(x[1] == y[0] for x, y in zip(b, b[1:]))
        assert all((hi - lo) % w == 0 for lo, hi in b[:-1])
        assert all(hi - lo >= min(w, batch) for lo, hi in b)
