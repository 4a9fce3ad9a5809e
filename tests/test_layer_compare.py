"""CPU tests of layer_compare: the reference's layer_N.npy / npc_cmp_print workflow (proj3/dnn_avx.py:15-28)."""
import numpy as np
import pytest

import layer_compare as lc


class _Node:
    def __init__(self, name, result):
        self.name, self.result = name, result


def _write(directory, tensors, base=0):
    directory.mkdir(exist_ok=True)
    for i, t in enumerate(tensors):
        np.save(directory / ("layer_%d.npy" % (base + i)), t)


def test_npc_cmp_print_walks_the_goldens_and_prints_like_the_reference(tmp_path, capsys):
    rng = np.random.default_rng(0)
    gold = [rng.standard_normal((1, 4, 4, 3)).astype(np.float32) for _ in range(3)]
    _write(tmp_path / "gold", gold)
    cmp = lc.NodeComparer(str(tmp_path / "gold"))
    mine = [gold[0].copy(), gold[1] + np.float32(0.5), gold[2].copy()]
    mine[2][0, 1, 2, 0] -= np.float32(0.25)
    rows = [cmp.npc_cmp_print(_Node("conv2d_%d" % i, m)) for i, m in enumerate(mine)]
    assert [r["index"] for r in rows] == [0, 1, 2] and cmp.npc == 3
    assert rows[0]["max_abs_diff"] == 0.0
    assert rows[1]["max_abs_diff"] == pytest.approx(0.5, abs=1e-6) and rows[2]["max_abs_diff"] == pytest.approx(0.25, abs=1e-6)
    out = capsys.readouterr().out.split("\n")
    assert out[0] == "conv2d_0" and float(out[1]) == 0.0 and out[2] == ""          # name, max |diff|, blank line
    assert out[3] == "conv2d_1" and float(out[4]) == pytest.approx(0.5, abs=1e-6)
    with pytest.raises(ValueError, match="golden"):
        lc.NodeComparer(str(tmp_path / "gold")).npc_cmp_print(_Node("x", np.zeros((2, 2), np.float32)))


def test_compare_dirs_handles_the_one_based_dumps_of_the_openblas_engine(tmp_path):
    rng = np.random.default_rng(1)
    tensors = [rng.standard_normal((2, 3)).astype(np.float32) for _ in range(4)]
    _write(tmp_path / "avx", tensors, base=0)            # proj3/dnn_avx.py:54-58: layer_0 ...
    _write(tmp_path / "openblas", tensors, base=1)       # proj3/dnn_openblas.py:47-50: layer_1 ...
    rows = lc.compare_dirs(str(tmp_path / "openblas"), str(tmp_path / "avx"), result_base=1, quiet=True,
                           names=["conv2d_0", "bias_add_0", "batch_norm_0", "leaky_relu_0"])
    assert len(rows) == 4 and all(r["max_abs_diff"] == 0.0 for r in rows) and rows[3]["name"] == "leaky_relu_0"
    rows = lc.compare_dirs(str(tmp_path / "openblas"), str(tmp_path / "avx"), result_base=1, golden_base=1, quiet=True)
    assert len(rows) == 3 and all(r["max_abs_diff"] > 0 for r in rows)                         # misaligned by one node
    with pytest.raises(FileNotFoundError):
        lc.compare_dirs(str(tmp_path / "nothing"), str(tmp_path / "avx"))
