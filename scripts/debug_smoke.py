"""Locate the first op that produces non-finite values in the tiny smoke model and compare every native
op against the torch reference on the same inputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quintnet_b200 import ops
from quintnet_b200.ops import reference, cuda_ops
from quintnet_b200.models.gpt2 import GPT2Config, GPT2Model

dev = torch.device("cuda")
cfg = GPT2Config(vocab_size=1024, n_positions=128, n_embd=256, n_layer=2, n_head=4)
model = GPT2Model(cfg, device=dev, dtype=torch.bfloat16, seed=0)
ids = torch.randint(0, cfg.vocab_size, (8, 128), device=dev)


def hook(name):
    def f(mod, inp, out):
        outs = out if isinstance(out, (tuple, list)) else (out,)
        for o in outs:
            if torch.is_tensor(o) and not torch.isfinite(o.float()).all():
                print(f"NON-FINITE after {name}: {type(mod).__name__}", flush=True)
    return f


for n, m in model.named_modules():
    m.register_forward_hook(hook(n))

# op-by-op checks on realistic shapes
def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-9)).item()

x = torch.randn(1024, 256, device=dev, dtype=torch.bfloat16)
for (N, K) in [(768, 256), (256, 256), (1024, 256), (256, 1024)]:
    a = torch.randn(1024, K, device=dev, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=dev, dtype=torch.bfloat16) * 0.05
    b = torch.randn(N, device=dev, dtype=torch.bfloat16)
    for bn in (128, 256):
        y = cuda_ops.gemm(a, w, 1024, N, K, epilogue=cuda_ops.EPI_BIAS, bias=b, block_n=bn)
        print(f"gemm 1024x{N}x{K} bn={bn} rel={rel(y, a.float() @ w.float().t() + b.float()):.2e} finite={torch.isfinite(y.float()).all().item()}")
    dy = torch.randn(1024, N, device=dev, dtype=torch.bfloat16)
    dx = cuda_ops.gemm(dy, w, 1024, K, N, b_mn=True)
    print(f"  dgrad rel={rel(dx, dy.float() @ w.float()):.2e}")
    dw = cuda_ops.gemm(dy, a, N, K, 1024, a_mn=True, b_mn=True)
    print(f"  wgrad rel={rel(dw, dy.float().t() @ a.float()):.2e}")

with torch.no_grad():
    emb = model.embedding(ids)
    print("embedding finite", torch.isfinite(emb.float()).all().item())
    ln = ops.layer_norm(emb, model.blocks[0].ln_1.weight, model.blocks[0].ln_1.bias, 1e-5)
    print("ln finite", torch.isfinite(ln.float()).all().item(), "rel", rel(ln, reference.layer_norm(emb.float(), model.blocks[0].ln_1.weight.float(), model.blocks[0].ln_1.bias.float())))
    d = ops.dropout_add(emb, emb, 0.1, True)
    print("dropout_add finite", torch.isfinite(d.float()).all().item())

ls, n = model(ids, labels=ids)
print("loss", (ls / n).item())
(ls / n).backward()
for name, p in model.named_parameters():
    if p.grad is not None and not torch.isfinite(p.grad.float()).all():
        print("NON-FINITE grad", name)
print("done")

# ---- tiny GPT-2 (head_dim 32 -> library attention) all-gradient comparison vs fp32 CPU
cfg2 = GPT2Config(vocab_size=512, n_positions=64, n_embd=128, n_layer=2, n_head=4, resid_pdrop=0.0, embd_pdrop=0.0, attn_pdrop=0.0)
ref = GPT2Model(cfg2, seed=3)
gpu = GPT2Model(cfg2, device=dev, dtype=torch.bfloat16, seed=3)
ids2 = torch.randint(0, 512, (8, 64))
lr_, nr = ref(ids2, labels=ids2); (lr_ / nr).backward()
lg, ng = gpu(ids2.cuda(), labels=ids2.cuda()); (lg / ng).backward()
print("tiny loss", (lr_ / nr).item(), (lg / ng).item())
for (n1, p1), (n2, p2) in zip(ref.named_parameters(), gpu.named_parameters()):
    print(f"  grad rel {n1:45s} {rel(p2.grad.cpu(), p1.grad):.3e}")
