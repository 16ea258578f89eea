"""TEST INFRASTRUCTURE -- torch restatement of the three glue kernels of quip_b200/csrc/glue.cu.

Each function restates, operation by operation, what the HF Llama modules compute between the linears of a decoder
layer -- the modules the reference's eval loop calls at llama.py:227 through `layer(...)`:
  rmsnorm   transformers modeling_llama.LlamaRMSNorm.forward, preceded by the decoder layer's `residual + hidden_states`
  rope_     modeling_llama.apply_rotary_pos_emb / rotate_half, on token-major (rows, heads*head_dim) storage, in place
  silu_mul  modeling_llama.LlamaMLP.forward: act_fn(gate_proj(x)) * up_proj(x) with act_fn = SiLU
Pinned by tests/test_fused_layer.py against the HF modules themselves (bit-exact on the CPU in fp16 and fp32).
Only tests/ and __graft_entry__.smoke() may import this module; the product (quip_b200/fused.py) takes it as an injected
`ops` object in the CPU tests and never imports it.
"""
import torch


class TorchGlue:
    def rmsnorm(self, x, weight, eps, residual=None):
        s = x if residual is None else x + residual                 # decoder layer: residual + hidden_states
        h = s.to(torch.float32)
        var = h.pow(2).mean(-1, keepdim=True)
        h = h * torch.rsqrt(var + eps)
        y = weight * h.to(x.dtype)
        return y if residual is None else (s, y)

    def rope_(self, q, k, cos, sin, head_dim):
        for t in (q, k):
            rows = t.numel() // t.shape[-1]
            v = t.view(rows, -1, head_dim)                           # (rows, heads, hd): cos/sin broadcast over heads
            c, s = cos.view(rows, 1, head_dim), sin.view(rows, 1, head_dim)
            half = head_dim // 2
            rot = torch.cat((-v[..., half:], v[..., :half]), dim=-1)
            v.copy_((v * c) + (rot * s))

    def silu_mul(self, gate, up):
        return torch.nn.functional.silu(gate) * up
