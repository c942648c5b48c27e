"""BASELINE.json configs[4] on ONE GPU (parity-style run, not a bench line): EmailVerifier with maxBodyLength = 16384
(10.2 M constraints, Groth16 domain 2^24), a small batch of synthetic emails with 12 KB bodies.
Checks: witness of email 0 == CPU oracle bit for bit, every proof verifies, a tampered email is rejected.
   python scripts/config5.py [batch]"""
import ctypes, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-email-verify_b200", "host")); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("ZKE_LANES", "4")
import zkemail_b200 as z
import zkutil

BATCH = int(sys.argv[1]) if len(sys.argv) > 1 else 8
t0 = time.perf_counter()
circuit = z.Circuit("EmailVerifier", [1024, 16384, 121, 17])
info = circuit.info
print("circuit: %d constraints, %d signals, %d levels, domain 2^%d (built in %.1f s)" % (info.n_constraints, info.n_vars, info.n_levels, info.domain_log2, time.perf_counter() - t0), flush=True)
key = z.synthetic.generate_key()
packed = []
for i in range(BATCH):
    em = z.synthetic.make_signed_email(i, key, body_len=12288)
    dk = z.verify_dkim_signature(em, resolver=lambda n, t: [z.synthetic.key_record(key)])
    packed.append(circuit.pack_inputs(z.generate_email_verifier_inputs_from_dkim_result(dk, {"maxBodyLength": 16384})))
t0 = time.perf_counter()
zk = z.Zkey(circuit, seed=77, device=0)
print("setup (toy key + H table): %.1f s" % (time.perf_counter() - t0), flush=True)
ctx = z.Context(circuit, zk, device=0, max_batch=BATCH)
blob = b"".join(packed)
t0 = time.perf_counter()
wt, status = ctx.witness(packed[0], 1)
print("witness of one email: %.3f s" % (time.perf_counter() - t0), flush=True)
rc = zkutil.ref_view(circuit)
wbuf = ctypes.create_string_buffer(32 * (info.n_vars + info.n_temps))
assert zkutil.ref.zkref_witness(ctypes.byref(rc), packed[0], wbuf) == 0
assert wt == wbuf.raw[: 32 * info.n_vars], "GPU witness differs from the CPU oracle"
print("witness == CPU oracle (%d signals, bit for bit)" % info.n_vars, flush=True)
ctx.fullprove(blob, BATCH)                        # warm-up
t0 = time.perf_counter()
proofs, publics, status = ctx.fullprove(blob, BATCH)
dt = time.perf_counter() - t0
print("fullprove batch %d: %.3f s => %.2f proofs/s" % (BATCH, dt, BATCH / dt), flush=True)
vkey = zk.vkey()
from oracle import bn254
npub = info.n_public
for e in (0, BATCH - 1):
    proof, pubs = z.proof_to_json(proofs[256 * e:256 * e + 256], publics[32 * npub * e:32 * npub * (e + 1)], npub)
    assert z.verify(vkey, pubs, proof), "product verifier rejected proof %d" % e
    assert bn254.groth16_verify(vkey, pubs, proof), "oracle verifier rejected proof %d" % e
print("proofs 0 and %d verify under the oracle's fixture-pinned verifier" % (BATCH - 1))
bad = bytearray(packed[0]); off = 32 * circuit.groups["emailBody"][0]
first, count, _ = circuit.groups["emailBody"]
base = 1 + info.n_outputs
bad[32 * (first - base + 100)] ^= 1
try:
    ctx.witness(bytes(bad), 1)
    raise SystemExit("tampered body was accepted")
except z.AssertFailed as ex:
    print("tampered body rejected:", str(ex)[:80])
ctx.profile(True); ctx.fullprove(blob, 1); prof = ctx.profile_get()
print(json.dumps({k: round(v["ms"] / max(1, v["count"]), 3) for k, v in prof.items()}))
