"""Can the UNMODIFIED reference models be captured in a CUDA graph with the drop-in installed
?  RAFTStereo and SMoEStereo (ViT-L + SMoE PEFT), 544x960, 32 iterations, default
mixed precision; eager vs graph replay, stock corr vs install()."""
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import smoe_stereo_b200 as sb  # noqa: E402
from baseline import ref_models as R  # noqa: E402

dev = torch.device("cuda", 0)
kind = sys.argv[1] if len(sys.argv) > 1 else "raft"
args = R.make_args("reg", mixed_precision=True, model="smoe" if kind == "smoe" else "raft")
model = R.build_model(kind, args, seed=0, device=dev)
left, right = R.synthetic_pair(1, 544, 960, max_disp=64.0, seed=5, device=dev)
sl, sr = left.clone(), right.clone()


def fwd():
    with torch.no_grad(), warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return model(sl, sr, iters=32, test_mode=True)["disp_preds"][-1]


def timed(fn, n=3):
    fn(); fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n, out


for arm in ("stock", "ours"):
    sb.uninstall()
    if arm == "ours":
        sb.install()
    t_eager, d_eager = timed(fwd)
    msg = f"{kind} {arm}: eager {1 / t_eager:6.2f} frames/s"
    try:
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        with torch.cuda.stream(side):
            fwd()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=side):
                static_out = fwd()
        torch.cuda.synchronize()
        t_graph, _ = timed(lambda: (g.replay(), static_out)[1])
        same = torch.equal(static_out, d_eager)
        msg += f"   CUDA graph {1 / t_graph:6.2f} frames/s  (replay equals eager: {same})"
    except Exception as e:      # noqa: BLE001
        msg += f"   CUDA graph capture failed: {type(e).__name__}: {str(e)[:160]}"
    print(msg, flush=True)
sb.uninstall()
