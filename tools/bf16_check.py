import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo")); sys.path.insert(0, os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "tests"))
import torch, numpy as np
import codetr_b200 as cb
from codetr_b200 import workloads as W
from parity import bf16_ulp_errors, max_rel
dev = torch.device("cuda:0")
KEYS = ("value", "spatial_shapes", "level_start_index", "sampling_loc", "attn_weight")
for name in ("swinl_enc_1152x768", "swinl_dec_1152x768", "swinl_enc_1920x1280"):
    wl = W.CONFIGS[name]
    inp = W.make_inputs(wl, batch=1)
    d = {k: torch.from_numpy(getattr(inp, k)) for k in KEYS}
    d = {k: (v.to(dev) if v.dtype == torch.int64 else v.to(device=dev, dtype=torch.bfloat16)) for k, v in d.items()}
    ref = cb.multi_scale_deformable_attention(d["value"].float(), d["spatial_shapes"], d["level_start_index"], d["sampling_loc"].float(), d["attn_weight"].float()).cpu().numpy()
    for fl, nm in ((cb.FLAG_MATH_EXACT, "exact"), (cb.FLAG_MATH_FHFMA, "fhfma"), (0, "default"), (0, "nosplit")):
        os.environ.pop("MSDA_B200_BF16_SPLIT", None)
        if nm == "nosplit":
            os.environ["MSDA_B200_BF16_SPLIT"] = "0"
        out = cb.multi_scale_deformable_attention(*(d[k] for k in KEYS), flags=fl).float().cpu().numpy()
        calls = [cb.PreparedForward(*(d[k] for k in KEYS), flags=fl)]
        for _ in range(20): calls[0]()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(200): calls[0]()
        e.record(); torch.cuda.synchronize()
        print(name, nm, "max_rel %.3e" % max_rel(out, ref), "ulp %.3f" % bf16_ulp_errors(out, ref), "rel_l2 %.3e" % (np.linalg.norm(out-ref)/np.linalg.norm(ref)), "%.2f us (L2-warm)" % (1e3*s.elapsed_time(e)/200), cb.last_variant())
