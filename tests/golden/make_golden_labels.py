"""Authoring-container script: golden training labels produced by the UNMODIFIED reference.

The reference builds its labels inline in `train_one_epoch` (open_flamingo/train/train_utils.py:102-106 for the
LAION batch, :126-149 for the interleaved MMC4 batch), so there is no function to call.  This script imports that
module from /root/reference as is and drives `train_one_epoch` on the CPU with stand-in collaborators (a model that
records the `labels` / `lang_x` it is called with, no-op optimizer / scheduler / wandb); the label tensors the
reference handed to the model are committed as tests/golden/labels.pt.  Nothing here is reachable from the product.

    python tests/golden/make_golden_labels.py        # needs /root/reference; writes tests/golden/labels.pt
"""
import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/open_flamingo/train/train_utils.py"
MEDIA, EOC, PAD, VOCAB = 57, 58, 59, 60


def load_reference_train_utils():
    try:
        import wandb  # noqa: F401
    except Exception:
        sys.modules["wandb"] = types.ModuleType("wandb")
    spec = importlib.util.spec_from_file_location("ref_train_utils", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class Tok:
    pad_token_id = PAD

    def __call__(self, text, add_special_tokens=False):
        return {"input_ids": [{"<image>": MEDIA, "<|endofchunk|>": EOC}[text]]}

    def batch_decode(self, ids):
        return [str(r.tolist()) for r in ids]


class Recorder(torch.nn.Module):
    """Stands in for DDP(Flamingo): records what the loop feeds it; its loss is linear in the LM input embedding so
    the embedding gradient the loop leaves behind (train_utils.py:172-194) is known in closed form."""

    def __init__(self, coeff=None):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(()))
        self.calls = []
        self.coeff = coeff
        if coeff is not None:
            emb = torch.nn.Embedding(*coeff.shape)
            lang = types.SimpleNamespace(get_input_embeddings=lambda: emb)
            self.emb = emb
            object.__setattr__(self, "module", types.SimpleNamespace(lang_encoder=lang))

    def forward(self, vision_x, lang_x, attention_mask, labels):
        self.calls.append({"lang_x": lang_x.clone(), "labels": labels.clone(), "vision_shape": tuple(vision_x.shape)})
        loss = self.w * 0.0 + 1.0
        if self.coeff is not None:
            loss = loss + (self.emb.weight * self.coeff).sum()
        return (loss,)


class Loader(list):
    @property
    def num_batches(self):
        return len(self)


class Noop:
    param_groups = [{"lr": 0.0}]

    def step(self, *a, **k):
        pass

    def zero_grad(self, *a, **k):
        pass

    def log(self, *a, **k):
        pass


def handmade_rows(T):
    """Edge cases of the interleaved rule; every row is padded / cut to T."""
    I, E, P = MEDIA, EOC, PAD
    rows = [
        [1, 2, 3, 4, 5, 6, 7, 8],                                   # no <image> at all
        [I, 1, 2, 3, E, I, 4, 5, E],                                # image first, chunk ends the row
        [1, 2, E, 3, I, 4, 5, E, 6, 7, I, 8, E, I, 9],              # text + eoc before the first image; eoc, text, image
        [I, 1, E, E, 2, I, 3],                                      # two eocs in a row
        [I, E, I, E, I, 1],                                         # empty chunks
        [I, I, 1, 2, E, 3, 4, 5],                                   # consecutive images; trailing text after eoc
        [P, P, I, 1, 2, E, P, P, I, 3],                             # pads before / between (left / middle padding)
        [I, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15],     # never closed
        [E, E, E, I, 1],                                            # only eocs before the image
    ]
    out = torch.full((len(rows), T), PAD, dtype=torch.int64)
    for r, row in enumerate(rows):
        row = row[:T]
        out[r, :len(row)] = torch.tensor(row)
    return out


def random_rows(B, T, seed, p_img=0.04, p_eoc=0.05, p_pad_tail=0.5):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, MEDIA, (B, T), generator=g)
    u = torch.rand((B, T), generator=g)
    ids[u < p_img] = MEDIA
    ids[(u >= p_img) & (u < p_img + p_eoc)] = EOC
    for b in range(B):
        if torch.rand((), generator=g) < p_pad_tail:
            n = int(torch.randint(1, T, (), generator=g))
            ids[b, T - n:] = PAD
    return ids


def main():
    ref = load_reference_train_utils()
    steps = []
    for T, seed in ((16, 0), (48, 1), (300, 2), (517, 3)):          # 300 / 517 cross the kernel's 256-token passes
        mmc4 = torch.cat([handmade_rows(T), random_rows(7, T, seed)], 0)
        laion = torch.cat([handmade_rows(T)[:4], random_rows(4, T, 100 + seed, p_eoc=0.0)], 0)
        steps.append((laion, mmc4))
    laion_loader = Loader()
    mmc4_loader = Loader()
    for laion, mmc4 in steps:
        laion_loader.append((torch.zeros(laion.shape[0], 3, 4, 4), (laion, torch.ones_like(laion))))
        mmc4_loader.append((torch.zeros(mmc4.shape[0], 2, 3, 4, 4),
                            [(row[None], torch.ones_like(row)[None]) for row in mmc4]))
    args = types.SimpleNamespace(
        num_epochs=1, precision="fp32", fsdp=False, fsdp_use_orig_params=False, rank=0, world_size=1,
        gradient_accumulation_steps=1, loss_multiplier_laion=1.0, loss_multiplier_mmc4=1.0,
        freeze_lm_embeddings=True, report_to_wandb=False, logging_steps=10 ** 9, batch_size_laion=8, batch_size_mmc4=16)
    model = Recorder()
    ref.train_one_epoch(args=args, model=model, epoch=0, laion_loader=laion_loader, mmc4_loader=mmc4_loader,
                        tokenizer=Tok(), optimizer=Noop(), lr_scheduler=Noop(), device_id="cpu", wandb=Noop())
    assert len(model.calls) == 2 * len(steps)
    cases = []
    for i, (laion, mmc4) in enumerate(steps):
        c_l, c_m = model.calls[2 * i], model.calls[2 * i + 1]
        assert torch.equal(c_l["lang_x"], laion) and torch.equal(c_m["lang_x"], mmc4)
        # ids < 60 and labels >= -100: int16 keeps the fixture small (tests widen to int64 again)
        cases.append({"interleaved": False, "input_ids": laion.to(torch.int16), "labels": c_l["labels"].to(torch.int16)})
        cases.append({"interleaved": True, "input_ids": mmc4.to(torch.int16), "labels": c_m["labels"].to(torch.int16)})
    # second pass with trainable LM embeddings: the loop masks their gradient down to the two added tokens
    coeff = torch.randn(VOCAB, 6, generator=torch.Generator().manual_seed(9))
    model2 = Recorder(coeff)
    args.freeze_lm_embeddings = False
    ref.train_one_epoch(args=args, model=model2, epoch=0, laion_loader=laion_loader, mmc4_loader=mmc4_loader,
                        tokenizer=Tok(), optimizer=Noop(), lr_scheduler=Noop(), device_id="cpu", wandb=Noop())
    out = {"meta": {"torch": str(torch.__version__), "reference": "open_flamingo/train/train_utils.py:94-194"},
           "media_id": MEDIA, "eoc_id": EOC, "pad_id": PAD, "cases": cases,
           "embed_coeff": coeff, "embed_backwards": len(model2.calls), "embed_grad_after": model2.emb.weight.grad.clone()}
    torch.save(out, os.path.join(HERE, "labels.pt"))
    n = sum(c["labels"].numel() for c in cases)
    print(f"wrote labels.pt: {len(cases)} cases, {n} labels")


if __name__ == "__main__":
    main()
