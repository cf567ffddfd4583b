"""Where does the gate-gradient error come from?  (round-2 experiment; prints a per-layer table)

For the OF-3B model at 4 x (2 images, 256 tokens): relative L2 error against the fp32 oracle of
  * the gradient arriving at every decoder position's output (hidden-state gradient), ours vs amp-oracle;
  * attn_gate / ff_gate gradients, ours vs amp-oracle;
and, for one gated block in isolation with an EXACT upstream gradient, the gate gradients recomputed in fp64 from our
own saved branch tensors (isolates the gate_bwd kernel from everything upstream).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import test_fullsize_parity_gpu as T  # noqa: E402

bf16 = torch.bfloat16


def hidden_grad_hooks(layers, store):
    hs = []
    for i, layer in enumerate(layers):
        def fwd_hook(mod, args, out, i=i):
            t = out[0] if isinstance(out, tuple) else out
            if t.requires_grad:
                t.register_hook(lambda g, i=i: store.__setitem__(i, g.detach().float().clone()))
        hs.append(layer.register_forward_hook(fwd_hook))
    return hs


def main():
    from open_flamingo_b200.testing import MPT_1B, build_flamingo
    model, _, tok = build_flamingo(T.VIT_L14, MPT_1B, cross_attn_every_n_layers=1, device="cuda", gate_init=1.0, seed=0)
    model.train()
    orc, sd, trainable = T._oracle_of(model, 1)
    batch = T._batch(tok, 4, 2, 256, MPT_1B["vocab_size"], seed=6, first_image_at=5)
    g_ours, g_ref, g_amp = {}, {}, {}
    hs = hidden_grad_hooks(list(model.lang_encoder._get_decoder_layers()), g_ours)
    ours = T._run_ours(model, batch)
    for h in hs:
        h.remove()
    hs = hidden_grad_hooks(orc.blocks, g_ref)
    ref = T._run_oracle(orc, sd, trainable, batch, amp=False)
    for h in hs:
        h.remove()
    hs = hidden_grad_hooks(orc.blocks, g_amp)
    amp = T._run_oracle(orc, sd, trainable, batch, amp=True)
    for h in hs:
        h.remove()
    print("layer | dHidden err ours / amp | attn_gate err ours / amp (ref value) | ff_gate err ours / amp (ref value)")
    for i in range(24):
        a, f = f"lang_encoder.gated_cross_attn_layers.{i}.attn_gate", f"lang_encoder.gated_cross_attn_layers.{i}.ff_gate"
        print(f"{i:2d} | {T._rel(g_ours[i], g_ref[i]):.3e} {T._rel(g_amp[i], g_ref[i]):.3e} | "
              f"{T._rel(ours[2][a], ref[2][a]):.3e} {T._rel(amp[2][a], ref[2][a]):.3e} ({ref[2][a].item():+.3e}) | "
              f"{T._rel(ours[2][f], ref[2][f]):.3e} {T._rel(amp[2][f], ref[2][f]):.3e} ({ref[2][f].item():+.3e})")
    # ---- one block in isolation, exact upstream gradient
    from open_flamingo_b200 import fused, ops
    from open_flamingo_b200 import _lib as L
    blk = model.lang_encoder.gated_cross_attn_layers[0]
    torch.manual_seed(3)
    R, D = 1024, 2048
    x1 = torch.randn(R, D, device="cuda")
    w = torch.randn(R, D, device="cuda")
    ff = blk.ff
    out, saved = fused._ffn_forward(x1, ff[0].weight, ff[0].bias, ff[1].weight, ff[3].weight, blk.ff_gate)
    branch = saved[5].double()
    t = torch.tanh(blk.ff_gate.detach().double())
    self_consistent = ((1 - t * t) * (w.double() * branch).sum()).item()
    dgate = torch.zeros(1, device="cuda")
    ops.gate_bwd(w, saved[5], blk.ff_gate.detach(), dgate)
    # fp32 oracle branch on the same input
    from oracle import flamingo_oracle as O
    sdl = {k[len("lang_encoder.gated_cross_attn_layers.0."):]: v.detach() for k, v in sd.items()
           if k.startswith("lang_encoder.gated_cross_attn_layers.0.")}
    with T._NoTF32():
        b_ref = O.feed_forward(x1, sdl, "ff").double()
        with torch.autocast("cuda", dtype=bf16):
            b_amp = O.feed_forward(x1, sdl, "ff").double()
    ref_gate = ((1 - t * t) * (w.double() * b_ref).sum()).item()
    amp_gate = ((1 - t * t) * (w.double() * b_amp).sum()).item()
    print(f"isolated FFN: kernel dgate {dgate.item():+.6e}; fp64 from our own branch {self_consistent:+.6e}; "
          f"fp32-oracle branch {ref_gate:+.6e}; amp-oracle branch {amp_gate:+.6e}")
    print(f"branch rel L2 err: ours {T._rel(branch, b_ref):.3e}, amp {T._rel(b_amp, b_ref):.3e}; "
          f"mean signed (ours - ref) {((branch - b_ref).mean() / b_ref.abs().mean()).item():+.3e}, "
          f"(amp - ref) {((b_amp - b_ref).mean() / b_ref.abs().mean()).item():+.3e}")


if __name__ == "__main__":
    main()
