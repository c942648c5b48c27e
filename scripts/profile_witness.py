"""calculateWitness + checkConstraints (zke_witness) for a batch of the default circuit between cudaProfilerStart / Stop:
the command ncu wraps for the witness kernel (thread-block clusters, native SHA-256 / FpMul / regex-seeding ops) and the
batched constraint check.   PROFILE_BATCH (default 64)
    ncu --profile-from-start off --set full --clock-control none -k regex:"witness_kernel|check_rows_kernel" -o out python scripts/profile_witness.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "zk-email-verify_b200", "host"))
import zkemail_b200 as z

batch = int(os.environ.get("PROFILE_BATCH", "64"))
circuit = z.Circuit("EmailVerifier", [1024, 1536, 121, 17])
ctx = z.Context(circuit, None, device=0, max_batch=batch)
key = z.synthetic.generate_key()
packed = []
for i in range(batch):
    em = z.synthetic.make_signed_email(i, key)
    dk = z.verify_dkim_signature(em, resolver=lambda n, t: [z.synthetic.key_record(key)])
    packed.append(circuit.pack_inputs(z.generate_email_verifier_inputs_from_dkim_result(dk)))
blob = b"".join(packed)
ctx.witness(blob, batch, want_witness=False)                # warm-up, outside the profiled range
rt = ctypes.CDLL("libcudart.so")
rt.cudaProfilerStart()
_, status = ctx.witness(blob, batch, want_witness=False)
rt.cudaProfilerStop()
assert status == [-1] * batch
print("profiled zke_witness for a batch of", batch)
