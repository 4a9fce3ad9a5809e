"""layer_compare -- the reference's only verification workflow, kept usable against this engine: per-node
tensors saved as `layer_N.npy` and compared node by node with the maximum absolute difference
(`npc_cmp_print`, proj3/dnn_avx.py:15-28; goldens written by proj1/__init__.py:37-42, dumps written by
proj3/dnn_avx.py:54-58 and, 1-based, by proj3/dnn_openblas.py:47-50).

    eng = dnn_b200.DnnInferenceEngine(graph, debug=True)     # node-by-node run, writes ./intermediate/layer_N.npy
    eng.run(x)
    layer_compare.compare_dirs("./intermediate", "../../proj1/intermediate")

    layer_compare.compare_blocks(eng_fused, golden_dir)      # a fused run only has block outputs: compares each
                                                             # block's last node with that node's golden file
"""
import os

import numpy as np


def layer_path(directory, index):
    return os.path.join(directory, "layer_{}.npy".format(index))


class NodeComparer:
    """Stateful twin of the reference's npc / npc_path / npc_cmp_print trio: every call compares the next
    node with the next golden file and prints what the reference prints (name, max |diff|, blank line)."""

    def __init__(self, golden_dir, start=0, quiet=False):
        self.golden_dir, self.npc, self.quiet = golden_dir, start, quiet
        self.rows = []

    def npc_path(self):
        ret = layer_path(self.golden_dir, self.npc)
        self.npc += 1
        return ret

    def npc_cmp_print(self, obj):
        path = self.npc_path()
        golden = np.load(path)
        result = np.asarray(obj.result)
        if golden.shape != result.shape:
            raise ValueError("%s: result %s vs golden %s (%s)" % (obj.name, result.shape, golden.shape, path))
        diff = float(np.abs(result.astype(np.float64) - golden.astype(np.float64)).max()) if result.size else 0.0
        scale = float(np.abs(golden).max()) if golden.size else 0.0
        row = {"index": self.npc - 1, "name": obj.name, "max_abs_diff": diff,
               "normalised": diff / scale if scale > 0 else diff}
        self.rows.append(row)
        if not self.quiet:
            print(obj.name)
            print(diff)
            print()
        return row


def compare_dirs(result_dir, golden_dir, names=None, result_base=0, golden_base=0, quiet=False):
    """Compares layer_{result_base + i}.npy of result_dir with layer_{golden_base + i}.npy of golden_dir for
    i = 0, 1, ... while both exist (the OpenBLAS engine's dumps are 1-based: result_base or golden_base = 1).
    Returns the rows [{index, name, max_abs_diff, normalised}]."""
    class _Obj:
        pass
    cmp = NodeComparer(golden_dir, start=golden_base, quiet=quiet)
    i = 0
    while os.path.exists(layer_path(result_dir, result_base + i)) and os.path.exists(layer_path(golden_dir, golden_base + i)):
        o = _Obj()
        o.name = names[i] if names is not None and i < len(names) else "layer_{}".format(result_base + i)
        o.result = np.load(layer_path(result_dir, result_base + i))
        cmp.npc_cmp_print(o)
        i += 1
    if i == 0:
        raise FileNotFoundError("no layer_N.npy pairs in %s / %s" % (result_dir, golden_dir))
    return cmp.rows


def compare_blocks(engine, golden_dir, golden_base=0, quiet=False):
    """After a fused run of dnn_b200.DnnInferenceEngine: the tensors inside a fused block never exist, so each
    block's output is compared with the golden file of the block's LAST node (node order = execution order
    without the input node, the numbering of proj1/__init__.py:37-42)."""
    class _Obj:
        pass
    order = [n for n in engine._ordered_nodes()][1:]
    index = {node: i for i, node in enumerate(order)}
    rows = []
    for node, result in engine.block_results().items():
        cmp = NodeComparer(golden_dir, start=golden_base + index[node], quiet=quiet)
        o = _Obj()
        o.name, o.result = node.name, result
        rows.append(cmp.npc_cmp_print(o))
    return rows
