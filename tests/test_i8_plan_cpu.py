"""Host-side planning of the int8 digit-product contraction (csrc/gemm_i8.cuh gi8_plan): for every padded feature count the
out-blocks tile the rows in 16-row units, fit the 512 TMEM columns (6 accumulators x rows), the resident W2 digit planes plus at
least two ring stages fit in shared memory, and the CTAs of all blocks add up to the SM count.  No GPU involved."""
import ctypes

import pytest

from gpfy_b200 import _lib

S, A_STAGE, MAXBLK = 6, 6 * 4096, 32


def _plan(Mp, sms=148):
    out = (ctypes.c_int * (5 + 3 * MAXBLK))()
    assert _lib.load().gpfy_debug_i8_plan(Mp, sms, out, len(out)) == 0
    if not out[0]:
        return None
    nblk = out[1]
    blocks = [(out[4 + 3 * i], out[5 + 3 * i], out[6 + 3 * i]) for i in range(nblk)]
    return {"nblk": nblk, "nsa": out[2], "smem": out[3], "blocks": blocks, "ctas": out[4 + 3 * nblk]}


@pytest.mark.parametrize("Mp", list(range(32, 1025, 32)))
def test_plan_invariants(Mp):
    p = _plan(Mp)
    if Mp > 864:
        assert p is None          # the W2 planes of even a 32-row block no longer leave two ring stages
        return
    assert p is not None, Mp
    rows_seen = 0
    for i, (row0, rows, cta0) in enumerate(p["blocks"]):
        assert row0 == rows_seen and rows % 16 == 0 and 16 <= rows <= 80
        assert S * rows <= 512                                   # TMEM columns
        assert cta0 == (0 if i == 0 else p["blocks"][i - 1][2] + (cta0 - p["blocks"][i - 1][2]))
        rows_seen += rows
    assert rows_seen == Mp
    ctas = [b[2] for b in p["blocks"]] + [p["ctas"]]
    assert all(ctas[i + 1] > ctas[i] for i in range(p["nblk"])) and p["ctas"] == 148
    assert 2 <= p["nsa"] <= 4
    maxrows = max(b[1] for b in p["blocks"])
    assert p["smem"] >= S * maxrows * Mp + p["nsa"] * A_STAGE and p["smem"] <= 227 * 1024


def test_headline_and_bench_shapes():
    assert [b[1] for b in _plan(288)["blocks"]] == [80, 80, 64, 64] and _plan(288)["nsa"] == 3      # cfg3
    assert [b[1] for b in _plan(352)["blocks"]] == [80, 80, 64, 64, 64] and _plan(352)["nsa"] == 2  # cfg4 (13 degrees)
    assert [b[1] for b in _plan(256)["blocks"]] == [64, 64, 64, 64]                                   # cfg5 (d = 32)
    assert max(b[1] for b in _plan(576)["blocks"]) == 48                                              # T = 64
    assert _plan(64, sms=1) is not None and _plan(288, sms=2) is None                                 # fewer SMs than blocks
