"""CPU check on the built library: the SASS of every hot kernel carries the Blackwell instructions the
design claims (B200_PROFILING.md's mnemonic table) -- tcgen05.mma = UTCHMMA (UTCHMMA.2CTA for the CTA-pair
kernel), tcgen05.ld = LDTM, TMA tensor loads = UTMALDG, bulk copies = UBLKCP -- and no legacy mma.sync
(HMMA / IMMA without the UTC prefix).  cuobjdump cross-reads sm_100a code without a GPU."""
import os
import re
import shutil
import subprocess

import pytest

import _lib

CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"


@pytest.fixture(scope="module")
def sass_by_kernel():
    if not os.path.exists(CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    text = subprocess.run([CUOBJDUMP, "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    kernels, name = {}, None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            kernels[name] = []
        elif name:
            kernels[name].append(line)
    assert kernels, "no SASS found in " + _lib.LIB_PATH
    return {k: "\n".join(v) for k, v in kernels.items()}


def _family(sass, key):
    sel = {k: v for k, v in sass.items() if key in k}
    assert sel, "no kernel named *%s* in the library" % key
    return sel


def test_built_for_sm_100a_only():
    out = subprocess.run([CUOBJDUMP, "-lelf", _lib.LIB_PATH], capture_output=True, text=True, timeout=120).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("key", ["conv_umma_kernel", "conv_pair_kernel", "conv_rowpair_kernel", "conv0_kernel"])
def test_tensor_core_kernels_use_tcgen05_and_tmem(sass_by_kernel, key):
    for name, sass in _family(sass_by_kernel, key).items():
        assert "UTCHMMA" in sass, name                       # tcgen05.mma
        assert "LDTM" in sass, name                          # tcgen05.ld from TMEM in the epilogue
        assert "UTCBAR" in sass, name                        # tcgen05.commit -> mbarrier
        assert not re.search(r"(?<!UTC)\b[HI]MMA\.", sass), name   # no mma.sync / wmma
        if key == "conv_pair_kernel":
            assert "UTCHMMA.2CTA" in sass, name              # cta_group::2


@pytest.mark.parametrize("key", ["conv_umma_kernel", "conv_pair_kernel", "conv_rowpair_kernel"])
def test_operands_arrive_by_tma(sass_by_kernel, key):
    for name, sass in _family(sass_by_kernel, key).items():
        assert "UTMALDG" in sass, name                       # cp.async.bulk.tensor loads


def test_conv0_streams_raw_pixels_with_bulk_copies(sass_by_kernel):
    for name, sass in _family(sass_by_kernel, "conv0_kernel").items():
        assert "UBLKCP" in sass, name                        # cp.async.bulk (12-byte pixels: no tensor map possible)


def test_staged_output_rows_leave_by_bulk_store(sass_by_kernel):
    """Block 8 (125 channels = 500-byte rows): the tile is staged in shared memory and written shared -> global by
    cp.async.bulk (UBLKCP.G.S), not by per-lane stores.  The staged-output instantiations are the ones whose mangled
    name ends in the MT = 1, STAGE_OUT = true, POOL1 = false, POOL_FIRST = false template arguments.  (The split-K
    partial tiles of the latency plan leave the same way, so other instantiations carry the instruction too.)"""
    staged = {k: v for k, v in _family(sass_by_kernel, "conv_umma_kernel").items() if "ELi1ELb1ELb0ELb0EEEv" in k}
    assert len(staged) == 2, sorted(staged)                  # TF32 and BF16-operand heads
    for name, sass in staged.items():
        assert "UBLKCP.G.S" in sass, name
