"""Diagnostics: per-iteration timing of the witness kernel (CTA 0), grouped by the kind of the iteration's heaviest op.
Run on a GPU box:  python scripts/witness_trace.py [batch]"""
import os, sys, struct
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-email-verify_b200", "host")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import zkemail_b200 as z
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
c = z.Circuit("EmailVerifier", [1024, 1536, 121, 17])
ctx = z.Context(c, None, device=0, max_batch=batch)
key = z.synthetic.generate_key()
packed = []
for i in range(batch):
    em = z.synthetic.make_signed_email(i, key)
    dk = z.verify_dkim_signature(em, resolver=lambda n, t: [z.synthetic.key_record(key)])
    packed.append(c.pack_inputs(z.generate_email_verifier_inputs_from_dkim_result(dk)))
blob = b"".join(packed)
ctx.witness(blob, batch, want_witness=False)
path = os.path.join(ROOT, "gpurun_out", "wtrace.bin")
os.makedirs(os.path.dirname(path), exist_ok=True)
os.environ["ZKE_WITNESS_TRACE"] = path
ctx.witness(blob, batch, want_witness=False)
del os.environ["ZKE_WITNESS_TRACE"]
raw = open(path, "rb").read()
n = len(raw) // (8 + 12)
clk = np.frombuffer(raw[: 8 * n], dtype=np.uint64).astype(np.int64)
info = np.frombuffer(raw[8 * n:], dtype=np.uint32).reshape(n, 3)
dt = np.diff(clk, prepend=clk[0])
code = info[:, 0] & 0xff
print("iterations", n, "total cycles", clk[-1] - clk[0], "=> ms at 1.965 GHz: %.1f" % ((clk[-1] - clk[0]) / 1.965e6))
for cd, name in [(0, "LIN"), (1, "QUAD"), (2, "SHRAND"), (3, "INVZ"), (4, "FPMUL")]:
    m = code == cd
    if m.any():
        print("%-7s iters %6d  cycles total %12d (%.1f ms)  mean %9.0f  median %9.0f  max %9d" % (
            name, m.sum(), dt[m].sum(), dt[m].sum() / 1.965e6, dt[m].mean(), np.median(dt[m]), dt[m].max()))
full = info[:, 1] == 512
print("full iterations:", full.sum(), "cycles %.1f ms" % (dt[full].sum() / 1.965e6), " partial:", (~full).sum(), "%.1f ms" % (dt[~full].sum() / 1.965e6))
q = (code == 1)
for lo, hi in [(0, 1024), (1024, 2048), (2048, 4096), (4096, 1 << 30)]:
    m = q & (info[:, 2] >= lo) & (info[:, 2] < hi)
    if m.any():
        print("QUAD iterations with %d <= terms < %d: %d, mean cycles %.0f" % (lo, hi, m.sum(), dt[m].mean()))
top = np.argsort(-dt)[:15]
for k in top:
    print("iter %6d code %d ops %4d terms %5d cycles %d" % (k, code[k], info[k, 1], info[k, 2], dt[k]))
