"""Byte-level weight-file fixtures for tests/test_weights_io.py, assembled WITHOUT weights_io's writers
(and without the pickle module): a darknet .weights file with the old 4-byte `seen` header (version 0.1),
one with the 8-byte header (0.2), and a Python-2 style protocol-2 pickle (str keys and raw array bytes as
Python-2 `str`, i.e. what proj3/yolov2tiny.py:15-23 expects to decode with encoding='latin1').

Tiny architecture: conv 3x3 3->4 with batch norm, conv 1x1 4->5 without.  Values are small integers / 8 so
they are exact in float32 and readable in a hex dump.     python tests/golden/make_weight_fixtures.py
"""
import os
import struct

HERE = os.path.dirname(os.path.abspath(__file__))
CHANNELS, KSIZES = [3, 4, 5], [3, 1]


def layer_values():
    """Darknet-order vectors of both layers as plain Python lists (floats exact in binary32)."""
    l0 = {
        "biases": [0.5, -0.25, 1.0, 0.125],
        "scales": [1.5, -0.75, 0.0, 2.0],                  # one negative, one zero scale
        "mean": [0.25, -0.5, 0.125, 0.0],
        "variance": [1.0, 0.25, 4.0, 0.0625],
        # filters [O=4][I=3][H=3][W=3]: value = (o*27 + i*9 + h*3 + w - 50) / 8
        "filters": [(o * 27 + i * 9 + h * 3 + w - 50) / 8.0 for o in range(4) for i in range(3) for h in range(3) for w in range(3)],
    }
    l1 = {
        "biases": [0.0, 0.5, -1.0, 2.0, -0.125],
        "filters": [(o * 4 + i - 9) / 4.0 for o in range(5) for i in range(4)],       # [O=5][I=4][1][1]
    }
    return l0, l1


def darknet_bytes(major, minor, seen):
    l0, l1 = layer_values()
    out = struct.pack("<3i", major, minor, 0)
    out += struct.pack("<q" if major * 10 + minor >= 2 else "<i", seen)
    for key in ("biases", "scales", "mean", "variance", "filters"):
        out += struct.pack("<%df" % len(l0[key]), *l0[key])
    for key in ("biases", "filters"):
        out += struct.pack("<%df" % len(l1[key]), *l1[key])
    return out


# ---- a minimal protocol-2 pickle assembler (opcodes as Python 2's cPickle writes them for numpy arrays)
def _short_str(s):
    b = s if isinstance(s, bytes) else s.encode("latin1")
    return (b"U" + struct.pack("<B", len(b)) + b) if len(b) < 256 else (b"T" + struct.pack("<i", len(b)) + b)


def _int(v):
    return b"K" + struct.pack("<B", v) if 0 <= v < 256 else b"J" + struct.pack("<i", v)


def _ndarray(shape, floats):
    raw = struct.pack("<%df" % len(floats), *floats)
    shp = b"(" + b"".join(_int(d) for d in shape) + b"t"
    return (b"cnumpy.core.multiarray\n_reconstruct\n" + b"cnumpy\nndarray\n" + _int(0) + b"\x85" + _short_str("b") + b"\x87R"
            + b"(" + _int(1) + shp
            + b"cnumpy\ndtype\n" + _short_str("f4") + _int(0) + _int(1) + b"\x87R"
            + b"(" + _int(3) + _short_str("<") + b"NNN" + _int(-1) + _int(-1) + _int(0) + b"tb"
            + b"\x89" + _short_str(raw) + b"tb")


def py2_pickle_bytes():
    """list of two dicts in the REFERENCE's convention and layout (kernel HWIO)."""
    l0, l1 = layer_values()

    def hwio(filters, o, i, k):
        return [filters[((oo * i + ii) * k + h) * k + w] for h in range(k) for w in range(k) for ii in range(i) for oo in range(o)]

    d0 = [("kernel", (3, 3, 3, 4), hwio(l0["filters"], 4, 3, 3)), ("biases", (4,), l0["biases"]),
          ("moving_mean", (4,), l0["mean"]), ("moving_variance", (4,), l0["variance"]), ("gamma", (4,), l0["scales"])]
    d1 = [("kernel", (1, 1, 4, 5), hwio(l1["filters"], 5, 4, 1)), ("biases", (5,), l1["biases"])]
    out = b"\x80\x02]("
    for d in (d0, d1):
        out += b"}("
        for key, shape, vals in d:
            out += _short_str(key) + _ndarray(shape, vals)
        out += b"u"
    return out + b"e."


if __name__ == "__main__":
    for name, blob in (("tiny_darknet_v01.weights", darknet_bytes(0, 1, 77)),
                       ("tiny_darknet_v02.weights", darknet_bytes(0, 2, 1234567890123)),
                       ("tiny_py2_protocol2.pickle", py2_pickle_bytes())):
        with open(os.path.join(HERE, name), "wb") as h:
            h.write(blob)
        print(name, len(blob), "bytes")
