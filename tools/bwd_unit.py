"""Unit checks of the backward GEMM kernel families against torch (GPU box only)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from wavenet_autoencoders_b200 import _lib
L = _lib.lib()
torch.manual_seed(0)
st = _lib.stream_ptr()
for (M, N, K) in [(256, 256, 7680), (256, 832, 61440), (128, 64, 640), (64, 320, 1000)]:
    Kp = K
    A = torch.randn(K, M, device="cuda").to(torch.bfloat16)
    B = torch.randn(K, N, device="cuda").to(torch.bfloat16)
    C = torch.zeros(M, N, device="cuda")
    _lib.check(L.wae_gemm_bf16_nt(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, st), "nt")
    torch.cuda.synchronize()
    ref = A.float().t() @ B.float()
    print(f"wgrad  M={M} N={N} K={K}: rel err {float((C - ref).abs().max() / ref.abs().max()):.3e}")
for (M, N, K) in [(7680, 256, 832), (1000, 128, 512), (16000, 64, 5120), (300, 256, 64)]:
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    W = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    _lib.check(L.wae_gemm_bf16_tn_bf16out(A.data_ptr(), W.data_ptr(), out.data_ptr(), M, N, K, 0.5, st), "tn")
    torch.cuda.synchronize()
    ref = (A.float() @ W.float().t()) * 0.5
    print(f"dgrad  M={M} N={N} K={K}: rel err {float((out.float() - ref).abs().max() / ref.abs().max()):.3e}")
