"""BASELINE.json configs[3] per-GPU share (parity-style run, not a bench line): the Proof-of-Twitter circuit
(`TwitterVerifier(1024, 1536, 121, 17)`: EmailVerifier + body regex + PackRegexReveal + public address), batch of 64
emails = one GPU's quarter of the 256-email, 4-GPU configuration (the path shards by proofs).
Checks: witness of email 0 == CPU oracle, proofs verify, public signals = [pubkeyHash, PackBytes(user name), address].
   python scripts/config4.py [batch]"""
import ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-email-verify_b200", "host")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import zkemail_b200 as z
import zkutil
from oracle import bn254

BATCH = int(sys.argv[1]) if len(sys.argv) > 1 else 64
circuit = z.Circuit("TwitterVerifier", [1024, 1536, 121, 17])
info = circuit.info
print("circuit: %d constraints, %d signals, domain 2^%d, %d public signals" % (info.n_constraints, info.n_vars, info.domain_log2, info.n_public), flush=True)
key = z.synthetic.generate_key()
packed, names = [], []
for i in range(BATCH):
    name = "user%04d" % i
    em = z.synthetic.make_signed_email(i, key, marker="This email was meant for @" + name)
    dk = z.verify_dkim_signature(em, resolver=lambda n, t: [z.synthetic.key_record(key)])
    packed.append(circuit.pack_inputs(z.generate_twitter_verifier_inputs_from_dkim_result(dk, 0x1234567890ABCDEF1234567890ABCDEF12345678 + i)))
    names.append(name)
zk = z.Zkey(circuit, seed=44, device=0)
ctx = z.Context(circuit, zk, device=0, max_batch=BATCH)
blob = b"".join(packed)
wt, _ = ctx.witness(packed[0], 1)
rc = zkutil.ref_view(circuit)
wbuf = ctypes.create_string_buffer(32 * (info.n_vars + info.n_temps))
assert zkutil.ref.zkref_witness(ctypes.byref(rc), packed[0], wbuf) == 0
assert wt == wbuf.raw[: 32 * info.n_vars], "GPU witness differs from the CPU oracle"
print("witness == CPU oracle (%d signals, bit for bit)" % info.n_vars, flush=True)
ctx.fullprove(blob, BATCH)
t0 = time.perf_counter()
n = 2
for _ in range(n):
    proofs, publics, status = ctx.fullprove(blob, BATCH)
dt = (time.perf_counter() - t0) / n
print("fullprove batch %d: %.3f s => %.2f proofs/s per GPU" % (BATCH, dt, BATCH / dt), flush=True)
vkey = zk.vkey()
for e in (0, BATCH - 1):
    proof, pubs = z.proof_to_json(proofs[256 * e:256 * e + 256], publics[96 * e:96 * e + 96], 3)
    assert int(pubs[1]) == int.from_bytes(names[e].encode(), "little") and int(pubs[2]) == 0x1234567890ABCDEF1234567890ABCDEF12345678 + e
    assert bn254.groth16_verify(vkey, pubs, proof) and z.verify(vkey, pubs, proof)
print("proofs 0 and %d verify; public signals = [pubkeyHash, PackBytes(user name), address]" % (BATCH - 1))
