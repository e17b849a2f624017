"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/umnn_b200.h declares,
validates descriptors, and its host-only entry points work without a GPU (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN_DIR, REPO
from umnn_b200 import _native, build as native_build


@pytest.fixture(scope="module")
def lib():
    native_build.build()
    return _native.lib()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(REPO, "include", "umnn_b200.h")).read()
    declared = set(re.findall(r"UMNN_API\s+[\w\s\*]+?\b(umnn_\w+)\s*\(", header))
    assert declared == set(_native.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.umnn_abi_version() == _native.UMNN_ABI_VERSION


def test_desc_struct_matches_header_layout():
    # int32 abi, int32 layout, int64 n_samples, 3 x int32, int32[9], 4 x int32  -> 80 bytes
    assert ctypes.sizeof(_native.Desc) == 80
    assert _native.Desc.n_samples.offset == 8 and _native.Desc.widths.offset == 28


def test_cc_tables_match_reference_bitwise(lib):
    g = np.load(os.path.join(GOLDEN_DIR, "cc_weights.npz"))
    for Q in (1, 2, 5, 20, 30, 50, 100, 200):
        t = np.zeros(Q + 1, np.float32)
        w = np.zeros(Q + 1, np.float32)
        assert lib.umnn_cc_tables(Q, t.ctypes.data, w.ctypes.data) == 0
        np.testing.assert_array_equal(t, g[f"t_{Q}"])
        assert np.max(np.abs(w - g[f"w_{Q}"])) <= np.spacing(np.float32(np.max(np.abs(w))))
    assert lib.umnn_cc_tables(0, t.ctypes.data, w.ctypes.data) == -2
    assert b"nb_steps" in lib.umnn_last_error()


def test_descriptor_validation_and_sizes(lib):
    d = _native.make_desc(_native.LAYOUT_STRIDED_D, 10, 6, 30, [31, 200, 200, 200, 1], _native.ACT_LEAKY_RELU,
                          _native.OUT_ELU_PLUS_1, 50, _native.PREC_FP32)
    assert lib.umnn_param_count(d) == 31 * 200 + 200 + 2 * (200 * 200 + 200) + 200 + 1 == 87001
    assert lib.umnn_packed_params_bytes(d) >= 87001 * 4
    bad = _native.make_desc(_native.LAYOUT_CONTIG, 10, 6, 30, [31, 200, 1], 0, 0, 50)
    assert lib.umnn_param_count(bad) == -1 and b"CONTIG" in lib.umnn_last_error()
    bad = _native.make_desc(_native.LAYOUT_STRIDED_D, 10, 6, 30, [30, 200, 1], 0, 0, 50)
    assert lib.umnn_packed_params_bytes(bad) == 0 and b"widths[0]" in lib.umnn_last_error()
    bad = _native.make_desc(_native.LAYOUT_STRIDED_D, 10, 6, 30, [31, 300, 1], 0, 0, 50)
    assert lib.umnn_param_count(bad) == -1
    bad = _native.make_desc(_native.LAYOUT_STRIDED_D, 10, 6, 30, [31, 200, 1], 0, 0, 50)
    bad.abi_version = 99
    assert lib.umnn_param_count(bad) == -1 and b"abi_version" in lib.umnn_last_error()
    with pytest.raises(_native.NativeError):
        _native.check(lib.umnn_pack_params(d, None, None, None))
    # zero samples is a valid no-op even without a device
    d0 = _native.make_desc(_native.LAYOUT_STRIDED_D, 0, 6, 30, [31, 200, 1], 1, 0, 50)
    assert lib.umnn_cc_forward(d0, None, None, None, None, None, None, None, None, None, None, 0, None) == 0


def test_precision_modes_sizes_and_auto_resolution(lib):
    """Packed-block and workspace sizes of the precision modes (host-only arithmetic, no device needed), and how
    UMNN_PREC_AUTO resolves: FP16X3 where the tensor-core kernel serves the shape, FP32 elsewhere."""
    def desc(prec, hidden=(200, 200, 200)):
        return _native.make_desc(_native.LAYOUT_STRIDED_D, 1000, 6, 30, [31] + list(hidden) + [1], _native.ACT_LEAKY_RELU,
                                 _native.OUT_ELU_PLUS_1, 50, prec)
    bf16 = lib.umnn_packed_params_bytes(desc(_native.PREC_BF16X3))
    fp16 = lib.umnn_packed_params_bytes(desc(_native.PREC_FP16X3))
    fp32 = lib.umnn_packed_params_bytes(desc(_native.PREC_FP32))
    assert bf16 > 0 and fp32 >= 87001 * 4
    # FP16X3 block = the BF16X3 block (forward + dgrad blobs, rounded up to 256 bytes), the fp16 forward blobs, and
    # the FP32 block of the guarded re-run
    assert fp16 > bf16 + fp32 and (fp16 - fp32 - (bf16 + 255) // 256 * 256) * 2 < bf16 * 1.1
    assert lib.umnn_packed_params_bytes(desc(_native.PREC_AUTO)) == fp16
    assert lib.umnn_workspace_bytes(desc(_native.PREC_AUTO), 0) == 256       # overflow flag of the guarded re-run
    assert lib.umnn_workspace_bytes(desc(_native.PREC_FP16X3), 0) == 256
    assert lib.umnn_workspace_bytes(desc(_native.PREC_BF16X3), 0) == 0
    assert lib.umnn_workspace_bytes(desc(_native.PREC_FP32), 0) == 0
    # FP16X3 backward: 256-byte flag block + the tensor-core panels (which the FP32 re-run borrows)
    assert lib.umnn_workspace_bytes(desc(_native.PREC_AUTO), 1) >= 256 + lib.umnn_workspace_bytes(desc(_native.PREC_BF16X3), 1) > 256
    # one hidden layer: no tensor-core kernel -> AUTO is the FP32 kernel, explicit tensor-core precisions are refused
    one = desc(_native.PREC_AUTO, hidden=(64,))
    assert lib.umnn_packed_params_bytes(one) == lib.umnn_packed_params_bytes(desc(_native.PREC_FP32, hidden=(64,)))
    assert lib.umnn_packed_params_bytes(desc(_native.PREC_FP16X3, hidden=(64,))) == 0
    assert b"tensor-core" in lib.umnn_last_error()
    bad = desc(7)
    assert lib.umnn_packed_params_bytes(bad) == 0 and b"precision" in lib.umnn_last_error()


def test_packed_layout_id_follows_the_nb_steps_threshold(lib):
    """UMNN_PREC_AUTO resolves through the shared-memory fit, which depends on nb_steps: [200]^3 is the FP32 layout
    for small Q and the tensor-core layout for larger Q.  The layout id (the key of the Python-side packed-parameter
    cache) must differ exactly where the packed block differs, so an eval-mode network whose Q changes across the
    threshold can never be launched on a block in the other format."""
    def desc(Q, prec=_native.PREC_AUTO):
        return _native.make_desc(_native.LAYOUT_STRIDED_D, 1000, 6, 30, [31, 200, 200, 200, 1], _native.ACT_LEAKY_RELU,
                                 _native.OUT_ELU_PLUS_1, Q, prec)
    sizes = {Q: lib.umnn_packed_params_bytes(desc(Q)) for Q in (5, 10, 20, 50, 100, 400)}
    ids = {Q: lib.umnn_packed_layout_id(desc(Q)) for Q in sizes}
    assert all(v != 0 for v in ids.values())
    fp32_bytes = lib.umnn_packed_params_bytes(desc(5, _native.PREC_FP32))
    assert sizes[5] == fp32_bytes and sizes[100] > fp32_bytes           # the threshold exists for this shape
    for a in sizes:
        for b in sizes:
            if sizes[a] != sizes[b]:
                assert ids[a] != ids[b]
    assert ids[50] == ids[100]                                          # same layout -> the block is shared
    assert ids[5] == lib.umnn_packed_layout_id(desc(5, _native.PREC_FP32))
    assert lib.umnn_packed_layout_id(desc(50, _native.PREC_BF16X3)) != ids[50]
    bad = desc(50)
    bad.nb_steps = 0
    assert lib.umnn_packed_layout_id(bad) == 0


@pytest.mark.parametrize("hidden,Q,narrow", [([100, 100, 100, 100], 50, 1), ([100, 50, 50, 50, 50], 50, 1), ([126, 126], 50, 1),
                                             ([127, 127], 50, 0), ([200, 200, 200], 50, 0), ([64, 64, 64], 1000, 1)])
def test_kernel_shape_selection_is_host_arithmetic(lib, monkeypatch, hidden, Q, narrow):
    """Which shape of the tensor-core forward serves a descriptor (two CTAs per SM when every padded width fits 128
    tensor-memory columns and the parameters + context fit half of the shared memory) -- no device involved."""
    monkeypatch.delenv("UMNN_B200_TC_NARROW", raising=False)
    d = _native.make_desc(_native.LAYOUT_STRIDED_D, 1000, 6, 30, [31] + hidden + [1], _native.ACT_LEAKY_RELU,
                          _native.OUT_ELU_PLUS_1, Q, _native.PREC_AUTO)
    flag = ctypes.c_int32(-1)
    assert lib.umnn_tc_forward_occupancy(d, 1, ctypes.byref(flag), None) == 0
    assert flag.value == narrow
    monkeypatch.setenv("UMNN_B200_TC_NARROW", "0")
    assert lib.umnn_tc_forward_occupancy(d, 1, ctypes.byref(flag), None) == 0 and flag.value == 0
    one = _native.make_desc(_native.LAYOUT_STRIDED_D, 10, 6, 30, [31, 64, 1], 1, 0, 50)
    assert lib.umnn_tc_forward_occupancy(one, 1, ctypes.byref(flag), None) == _native_unsupported


_native_unsupported = -4


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", os.path.join(tmp_path, "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU/eager fallback"):
        _native.lib()


def test_backward_workspace_follows_the_chunk_plan(lib, monkeypatch):
    """Host-only view of the tensor-core backward's chunk plan (cc_backward_tc.cu: make_bwd_plan) through the workspace
    size it asks for: chunks are balanced (a call a little above one maximum chunk is cut into two equal chunks, not a full
    one and a remainder), UMNN_B200_BWD_TILES bounds the chunk, and from 606 K rows on the panels are hi-only except the two
    head panels (the all-hi measurement mode asks for less, hi + lo everywhere for more)."""
    def ws(B, **env):
        for k in ("UMNN_B200_BWD_TILES", "UMNN_B200_BWD_PANELS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        d = _native.make_desc(_native.LAYOUT_STRIDED_D, B, 6, 30, [31, 200, 200, 200, 1], _native.ACT_LEAKY_RELU,
                              _native.OUT_ELU_PLUS_1, 50, _native.PREC_BF16X3)        # rows = B * 6 * 53
        return int(lib.umnn_workspace_bytes(d, 1))

    full96 = ws(60000)                                  # 19 M rows: chunks of the maximum size (96 tiles x 148 CTAs = 1.82 M rows)
    assert ws(120000) == full96                         # the workspace is one chunk, however many chunks follow
    assert 0.9 * full96 < 3 * ws(60000, UMNN_B200_BWD_TILES="32") < 1.1 * full96
    # 1.1 maximum chunks' worth of rows -> two chunks of 0.55 each, not 1 + 0.1
    one_and_a_bit = ws(6300)                            # 2.0 M rows
    assert 0.5 * full96 < one_and_a_bit < 0.65 * full96
    # panel modes on a call above the hi-only threshold
    auto, all_hi, hilo = ws(10000), ws(10000, UMNN_B200_BWD_PANELS="hi"), ws(10000, UMNN_B200_BWD_PANELS="hilo")
    assert ws(10000, UMNN_B200_BWD_PANELS="hi_head") == auto
    assert all_hi < auto < hilo and auto < 1.25 * all_hi and hilo > 1.7 * all_hi
    # below the threshold every panel keeps its lo part
    assert ws(1000) == ws(1000, UMNN_B200_BWD_PANELS="hilo")
