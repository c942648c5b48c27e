"""One proof of the default circuit on a single lane between cudaProfilerStart / Stop: the command ncu wraps to produce
the per-kernel launch lists under profiles/ (see profiles/README.md).
    ZKE_LANES=1 ncu --profile-from-start off --metrics <...> --clock-control none --csv --log-file out.csv python scripts/profile_one.py
Environment: PROFILE_CIRCUIT="EmailVerifier:1024,1536,121,17" (default), PROFILE_SEED."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zk-email-verify_b200", "host"))
import zkemail_b200 as z

name, _, params = os.environ.get("PROFILE_CIRCUIT", "EmailVerifier:1024,1536,121,17").partition(":")
params = [int(p) for p in params.split(",")] if params else []
circuit = z.Circuit(name, params)
zk = z.Zkey(circuit, seed=int(os.environ.get("PROFILE_SEED", "20260923")), device=0)
ctx = z.Context(circuit, zk, device=0, max_batch=1)
key = z.synthetic.generate_key()
email = z.synthetic.make_signed_email(0, key)
dk = z.verify_dkim_signature(email, resolver=lambda n, t: [z.synthetic.key_record(key)])
opts = {"maxHeadersLength": params[0], "maxBodyLength": params[1]} if name == "EmailVerifier" else {}
packed = circuit.pack_inputs(z.generate_email_verifier_inputs_from_dkim_result(dk, opts))
ctx.fullprove(packed, 1)                                   # warm-up, outside the profiled range
rt = ctypes.CDLL("libcudart.so")
rt.cudaProfilerStart()
proofs, publics, status = ctx.fullprove(packed, 1)
rt.cudaProfilerStop()
proof, pubs = z.proof_to_json(proofs, publics, circuit.info.n_public)
assert status == [-1] and z.verify(zk.vkey(), pubs, proof)
print("profiled one proof; kernels launched so far:", z._lib.zke_kernel_launches())
