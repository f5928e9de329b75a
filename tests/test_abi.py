"""CPU: the C-ABI shared library loads and exports every symbol include/pq3d_b200.h declares
(no compute calls without a GPU), and bad arguments come back as error codes, not crashes."""
import ctypes

import pytest

from pq3d_b200 import _lib, build


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.lib()


def test_exports_match_header(lib):
    declared = _lib.declared_symbols()
    assert len(declared) >= 9
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/pq3d_b200.h but not exported"
    for name in _lib.SIGNATURES:
        assert name in declared, f"{name} bound in _lib.py but missing from the header"


def test_abi_version(lib):
    assert lib.pq3d_abi_version() == 1


def test_bad_arguments_return_error_codes(lib):
    rc = lib.pq3d_linear_bf16(None, 0, 0, 0, None, 0, 0, 0, None, 0, 0, 0, None, 0, 0, None, 0, 1, 1, 1, 1, 1.0, 0, 0,
                              0, None)
    assert rc == -1 and b"null operand" in lib.pq3d_last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    rc = lib.pq3d_linear_bf16(p, 8, 1, 0, p, 8, 1, 0, p, 8, 0, 0, None, 0, 0, None, 0, 1, 1, 7, 1, 1.0, 0, 0, 0, None)
    assert rc == -1 and b"multiple of 64" in lib.pq3d_last_error()
    rc = lib.pq3d_pack_mask(None, None, 0, 0, 0, None, None, 0, None)
    assert rc == -1


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libpq3d_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.lib()


def test_product_modules_refuse_cpu_tensors():
    """No CPU path in the product: without the test-only kernel emulation (tests/_cpu_ops.py, which monkeypatches
    pq3d_b200.ops inside a context manager) the decoder, in eval and in training, raises on CPU tensors."""
    import torch
    from pq3d_b200 import synth
    from pq3d_b200.query_encoder import QueryMaskEncoder
    w = synth.Workload("t", 1, 8, 16, ["pc"], "parallel", num_layers=1)
    enc = QueryMaskEncoder(None, **w.decoder_kwargs())
    inp, pw, _ = synth.make_decoder_inputs(w)
    enc.eval()
    with torch.no_grad(), pytest.raises((TypeError, ValueError), match="CUDA"):
        enc(synth.clone_input_dict(inp), pw)
    enc.train()
    with pytest.raises((TypeError, ValueError), match="CUDA"):
        enc(synth.clone_input_dict(inp), pw)


def test_scheduling_context_managers_restore_state():
    """ops.launch_priority / ops.shared_sm are plain host state around the launches (no compute): nesting and
    exceptions restore the previous value; without a GPU the C side clamps every priority to 0 and still returns OK."""
    from pq3d_b200 import ops
    assert ops._PRIORITY[0] == 0 and ops._SHARED_SM[0] is False
    with ops.launch_priority(-2):
        assert ops._PRIORITY[0] == -2
        with ops.launch_priority(0):
            assert ops._PRIORITY[0] == 0
        assert ops._PRIORITY[0] == -2
    assert ops._PRIORITY[0] == 0
    try:
        with ops.shared_sm(True), ops.launch_priority(-1):
            assert ops._SHARED_SM[0] is True
            raise KeyError("x")
    except KeyError:
        pass
    assert ops._PRIORITY[0] == 0 and ops._SHARED_SM[0] is False
