"""Experiment driver (GPU box): one context, several engine settings, proofs/s for each.
   python scripts/sweep.py "lanes=8,split=1" "lanes=16,split=1" ...   (env knobs: any ZKE_* name=value too)"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "zk-email-verify_b200", "host")); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("ZKE_LANES", "8")
import torch
import zkemail_b200 as z
import bench
BATCH = int(os.environ.get("SWEEP_BATCH", "64"))
circuit = z.Circuit(*bench.CIRCUIT)
key = z.synthetic.generate_key()
zk = z.Zkey(circuit, seed=bench.KEY_SEED, device=0)
ctx = z.Context(circuit, zk, device=0, max_batch=BATCH)
packed = b"".join(bench.make_inputs(z, circuit, BATCH, key))
ctx.upload_inputs(packed, BATCH)
L = z._lib
proofs = ctypes.create_string_buffer(256 * BATCH)
pubs = ctypes.create_string_buffer(32 * circuit.info.n_public * BATCH)
status = (ctypes.c_int32 * BATCH)()
err = ctypes.create_string_buffer(4096)
def step():
    rc = L.zke_fullprove(ctx.handle, None, BATCH, None, proofs, pubs, status, err, 4096)
    assert rc == 0, err.value
for spec in sys.argv[1:] or ["lanes=8,split=1"]:
    kv = dict(x.split("=") for x in spec.split(","))
    for k, v in kv.items():
        if k == "lanes": L.zke_ctx_set_lanes(ctx.handle, int(v))
        elif k == "split": os.environ["ZKE_SPLIT_STREAMS"] = v
        else: os.environ[k] = v
    step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 3
    for _ in range(n): step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    print("%-40s %.1f ms/step  %.2f proofs/s" % (spec, dt * 1e3, BATCH / dt), flush=True)
    if os.environ.get("SWEEP_PROFILE_EACH"):
        ctx.profile(True); step(); prof = ctx.profile_get(); ctx.profile(False)
        print("   ", {k: round(v["ms"] / max(1, v["count"]), 3) for k, v in prof.items()}, flush=True)
ctx.profile(True); step(); prof = ctx.profile_get(); ctx.profile(False)
print({k: round(v["ms"] / max(1, v["count"]), 3) for k, v in prof.items()})
