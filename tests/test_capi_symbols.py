"""The C-ABI library loads without a GPU and exports every function include/zkemail_b200.h declares."""
import os
import re
import ctypes

import zkemail_b200 as z
from zkemail_b200 import _lib as L

HEADER = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "zkemail_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(zke_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    names = _declared()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(L.lib, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_python_binding_covers_the_header():
    for n in _declared():
        assert hasattr(L, n), f"{n} has no ctypes signature in _lib.py"


def test_no_cpu_fallback_without_device():
    c = z.Circuit("Multiplier")
    if z.device_count() == 0:
        import pytest
        with pytest.raises(L.ZkeError, match="no CUDA device"):
            z.Zkey(c)
        with pytest.raises(L.ZkeError, match="no CUDA device"):
            z.Context(c)


def test_unknown_template_and_bad_inputs():
    import pytest
    with pytest.raises(L.ZkeError, match="unknown template"):
        z.Circuit("NoSuchTemplate")
    with pytest.raises(L.ZkeError, match="parameter asserts failed"):
        z.Circuit("EmailVerifier", [1000, 1536, 121, 17])       # maxHeadersLength % 64 != 0 (email-verifier.circom:43)
    c = z.Circuit("Multiplier")
    with pytest.raises(L.ZkeError, match="Signal not found"):
        c.pack_inputs({"a": 1, "b": 2, "zzz": 3})
    with pytest.raises(L.ZkeError, match="Not all inputs have been set"):
        c.pack_inputs({"a": 1})
    with pytest.raises(L.ZkeError, match="Too many values"):
        c.pack_inputs({"a": [1, 2], "b": 2})


def test_json_input_packing_matches_python():
    import json
    c = z.Circuit("FpMul", [2, 4])
    inp = {"a": ["1", "0", "1", "0"], "b": [0, 1, 1, 0], "p": ["1", "1", "1", "1"]}
    out = ctypes.create_string_buffer(32 * c.n_inputs)
    err = ctypes.create_string_buffer(512)
    assert L.zke_pack_inputs_json(c.handle, json.dumps(inp).encode(), out, len(out), err, 512) == 0, err.value
    assert out.raw == c.pack_inputs(inp)
    assert L.zke_pack_inputs_json(c.handle, b'{"a": [1,0,1,0]}', out, len(out), err, 512) != 0
    assert b"Not all inputs have been set" in err.value
