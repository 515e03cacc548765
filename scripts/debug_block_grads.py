"""GPU debugging aid: compare intermediate gradients of one bf16 MultiScaleBlock (dim change) with the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "portrait-mode-video_b200"))
import torch, torch.nn.functional as F
from oracle import detgen, mvit_oracle as orc
from pmv_b200 import functional as Fn, ops
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
T = torch.bfloat16 if (len(sys.argv) < 2 or sys.argv[1] == "bf16") else torch.float32
thw, sq, skv, heads = [4, 16, 16], [1, 2, 2], [1, 4, 4], 2
shapes = orc.block_param_shapes("", 96, 192, heads, thw, sq, skv)
P = {k: v.cuda() for k, v in detgen.det_params(shapes, 11).items()}
x0 = detgen.det_normal((2, 1 + 4 * 16 * 16, 96), 11, "x").cuda()

def nerr(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

G = {}
def keep(name, store):
    def hook(g):
        store[name] = g.detach().float().clone()
    return hook

# ---- oracle with hooks
go = {}
po = {k: v.clone().requires_grad_(True) for k, v in P.items()}
xo = x0.clone().requires_grad_(True)
xn = orc.layer_norm(xo, po["norm1.weight"], po["norm1.bias"]); xn.register_hook(keep("dxn", go))
B, N, _ = xn.shape
qkv = F.linear(xn, po["attn.qkv.weight"], po["attn.qkv.bias"]); qkv.register_hook(keep("dqkv", go))
q5 = qkv.reshape(B, N, 3, heads, 96).permute(2, 0, 3, 1, 4)
q, qs = orc.conv_pool_tokens(q5[0], thw, po["attn.pool_q.weight"], sq, True, po["attn.norm_q.weight"], po["attn.norm_q.bias"])
k, ks = orc.conv_pool_tokens(q5[1], thw, po["attn.pool_k.weight"], skv, True, po["attn.norm_k.weight"], po["attn.norm_k.bias"])
v, _ = orc.conv_pool_tokens(q5[2], thw, po["attn.pool_v.weight"], skv, True, po["attn.norm_v.weight"], po["attn.norm_v.bias"])
q.register_hook(keep("dq", go)); k.register_hook(keep("dk", go)); v.register_hook(keep("dv", go))
attn = (q * 96 ** -0.5) @ k.transpose(-2, -1)
attn = orc.add_rel_pos_bias(attn, q, True, qs, ks, po["attn.rel_pos_h"], po["attn.rel_pos_w"], po["attn.rel_pos_t"]).softmax(-1)
o = attn @ v
o = torch.cat([o[:, :, :1], o[:, :, 1:] + q[:, :, 1:]], 2).transpose(1, 2).reshape(B, -1, 192); o.register_hook(keep("do", go))
xb = F.linear(o, po["attn.proj.weight"], po["attn.proj.bias"])
xs = F.linear(xn, po["proj.weight"], po["proj.bias"]); xs.register_hook(keep("dxs", go))
xr, _ = orc.max_pool_tokens(xs, thw, [1, 3, 3], sq, True)
x1 = xr + xb; x1.register_hook(keep("dx1", go))
xn2 = orc.layer_norm(x1, po["norm2.weight"], po["norm2.bias"]); xn2.register_hook(keep("dxn2", go))
y = x1 + orc.mlp(xn2, po, "mlp.")
dy = detgen.det_normal(tuple(y.shape), 11, "dy").cuda()
y.backward(dy)
go["dx"] = xo.grad

# ---- ours with hooks
gm = {}
pm = {k: v.clone().requires_grad_(True) for k, v in P.items()}
x = x0.clone().requires_grad_(True)
xn_ = Fn.layer_norm(x, pm["norm1.weight"], pm["norm1.bias"], T); xn_.register_hook(keep("dxn", gm))
qkv_ = Fn.linear(xn_, pm["attn.qkv.weight"], pm["attn.qkv.bias"]); qkv_.register_hook(keep("dqkv", gm))
o_ = Fn.pool_attention(qkv_, pm["attn.pool_q.weight"], pm["attn.pool_k.weight"], pm["attn.pool_v.weight"],
                       pm["attn.norm_q.weight"], pm["attn.norm_q.bias"], pm["attn.norm_k.weight"], pm["attn.norm_k.bias"],
                       pm["attn.norm_v.weight"], pm["attn.norm_v.bias"], pm["attn.rel_pos_h"], pm["attn.rel_pos_w"], pm["attn.rel_pos_t"],
                       heads, thw, 2, 4, 96 ** -0.5, True, os.environ.get("PMV_TC_ATTENTION", "1") == "1")
o_.register_hook(keep("do", gm))
xs_ = Fn.linear(xn_, pm["proj.weight"], pm["proj.bias"], out_fp32=True); xs_.register_hook(keep("dxs", gm))
xr_ = Fn.maxpool_skip(xs_, thw)
x1_ = Fn.linear(o_, pm["attn.proj.weight"], pm["attn.proj.bias"], residual=xr_, out_fp32=True); x1_.register_hook(keep("dx1", gm))
xn2_ = Fn.layer_norm(x1_, pm["norm2.weight"], pm["norm2.bias"], T); xn2_.register_hook(keep("dxn2", gm))
y_ = Fn.mlp(xn2_, pm["mlp.fc1.weight"], pm["mlp.fc1.bias"], pm["mlp.fc2.weight"], pm["mlp.fc2.bias"], residual=x1_)
print("fwd y err", nerr(y_.detach(), y.detach()), " o err", nerr(o_.detach().float(), o.detach()), " x1", nerr(x1_.detach(), x1.detach()))
y_.backward(dy)
gm["dx"] = x.grad
for k_ in ("dxn2", "dx1", "do", "dxs", "dqkv", "dxn", "dx"):
    a, b = gm[k_].reshape(go[k_].shape).double(), go[k_].double()
    l2 = float((a - b).norm() / b.norm())
    frac = float(((a - b).abs() > 0.05 * b.abs().max()).double().mean())
    print(f"{k_:6s} max-norm err {nerr(a, b):.3e}  rel-L2 {l2:.3e}  frac(|err|>5% of max) {frac:.2e}  |ref|max {float(b.abs().max()):.3e}")
for k_ in sorted(pm):
    if k_.endswith("norm_k.bias"): continue
    print(f"  d{k_:28s} {nerr(pm[k_].grad, po[k_].grad):.3e}")
