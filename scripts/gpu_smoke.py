"""Stage timings of the full EmailVerifier(1024,1536) pipeline on one GPU (diagnostic, not the bench)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-email-verify_b200", "host"))
import zkemail_b200 as z

def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    t = time.time(); c = z.Circuit("EmailVerifier", [1024, 1536, 121, 17]); print("circuit build %.2fs" % (time.time() - t), c.info.n_vars, c.info.n_constraints, c.info.domain_log2, flush=True)
    key = z.synthetic.generate_key()
    packed = b""
    for i in range(batch):
        em = z.synthetic.make_signed_email(i, key)
        dk = z.verify_dkim_signature(em, resolver=lambda n, t: [z.synthetic.key_record(key)])
        packed += c.pack_inputs(z.generate_email_verifier_inputs_from_dkim_result(dk))
    t = time.time(); zk = z.Zkey(c, seed=1); print("setup %.2fs" % (time.time() - t), flush=True)
    t = time.time(); ctx = z.Context(c, zk, max_batch=batch); print("ctx open %.2fs" % (time.time() - t), flush=True)
    for rep in range(2):
        t = time.time(); _, st = ctx.witness(packed, batch, want_witness=False); dt = time.time() - t
        print("witness+check batch=%d: %.3fs (%.1f ms/email) status=%s" % (batch, dt, 1e3 * dt / batch, st[:4]), flush=True)
    for rep in range(2):
        t = time.time(); proofs, pubs, st = ctx.prove(batch); dt = time.time() - t
        print("prove batch=%d: %.3fs (%.1f ms/proof)" % (batch, dt, 1e3 * dt / batch), flush=True)
    vkey = zk.vkey()
    npub = c.info.n_public
    ok = all(z.verify(vkey, *reversed(z.proof_to_json(proofs[256 * k:256 * k + 256], pubs[32 * npub * k:32 * npub * (k + 1)], npub))) for k in range(batch))
    print("all proofs verify:", ok, "launches:", z._lib.zke_kernel_launches(), flush=True)

main()
