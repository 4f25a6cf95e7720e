"""TEST / BENCH INFRASTRUCTURE — the reference's torch-CUDA path as the speed comparator of BASELINE.json's
">= 8x the reference torch-CUDA path" target, measured IN the bench run on the same GPU.

/root/reference cannot travel to the GPU box, so this drives the oracle port (bit-identical arithmetic to the
reference, tests/test_oracle.py) on device="cuda" the way the reference's own benchmark drives the model
(moondream/torch/sample.py:159-207): batch-1, sequential over the images, one `.item()` per token
(moondream.py:482), eager first, then with the reference's `compile()` recipe (moondream.py:194-204):
`_vis_enc` and `_prefill` under torch.compile(fullgraph=True), `_decode_one_tok` under
torch.compile(fullgraph=True, mode="reduce-overhead"), warm-up runs before the timed ones.

    python -m oracle.torch_cuda_comparator --images 8 --tokens 64 [--compile] [--model moondream-2b]

Prints ONE JSON line.  Never imported by the product package.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle.moondream_oracle import Encoded, OracleModel  # noqa: E402


class SeamOracle(OracleModel):
    """OracleModel with the reference's three compile seams made explicit (same arithmetic)."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self._vis_enc = self.vision_encoder
        self._prefill = self.text_decoder
        self._decode_one_tok = self._decode_impl

    def _decode_impl(self, x, mask, pos_ids):
        """moondream.py:183-192."""
        hidden = self.text_decoder(x, mask, pos_ids)
        return self.lm_head(hidden), hidden

    def compile(self, backend: str = "inductor"):
        """moondream.py:194-204.  (`backend` other than inductor only to check graph capture on a box without a GPU.)
        The reference's weights are nn.Parameters and its KV caches module buffers, i.e. static addresses for
        Inductor's CUDA graphs; this port holds them in plain containers, so they are marked static explicitly —
        otherwise "reduce-overhead" either skips the graphs (mutated cache inputs) or copies every weight per replay."""
        for t in list(self.w.values()) + self.k_cache + self.v_cache + [self.rope, self.attn_mask]:
            torch._dynamo.mark_static_address(t)
        self._vis_enc = torch.compile(self.vision_encoder, fullgraph=True, backend=backend)
        self._prefill = torch.compile(self.text_decoder, fullgraph=True, backend=backend)
        if backend == "inductor":
            self._decode_one_tok = torch.compile(self._decode_impl, fullgraph=True, mode="reduce-overhead")
        else:
            self._decode_one_tok = torch.compile(self._decode_impl, fullgraph=True, backend=backend)

    def encode(self, image) -> Encoded:
        """moondream.py:206-268 through the seams."""
        from oracle.moondream_oracle import stitch_crops

        v = self.cfg.vision
        with torch.inference_mode():
            crops, tiling = self.prepare_crops(image)
            torch._dynamo.mark_dynamic(crops, 0)                    # moondream.py:209
            feats = self._vis_enc(crops)
            g = v.crop_size // v.enc_patch_size
            stitched = stitch_crops(feats[1:].view(-1, g, g, v.enc_dim), tiling, v.overlap_margin)
            img_emb = self.vision_projection(feats[0], stitched)
            bos = self.embed(torch.tensor([[self.cfg.tokenizer.bos_id]], device=self.device))
            x = torch.cat([bos, img_emb[None]], dim=1)
            n = x.size(1)
            self._prefill(x, self.attn_mask[:, :, 0:n, :], torch.arange(n, dtype=torch.long, device=self.device))
            return Encoded(n, [(self.k_cache[i][:, :, :n, :].clone(), self.v_cache[i][:, :, :n, :].clone())
                               for i in range(self.cfg.text.n_layers)])

    def answer(self, enc: Encoded, prompt, max_tokens: int):
        """_prefill_prompt + the generator of _generate_answer (moondream.py:280-321, 470-530), greedy."""
        tk = self.cfg.tokenizer
        out = []
        with torch.inference_mode():
            self.load_encoded(enc)
            x = self.embed(torch.tensor([list(prompt)], device=self.device))
            torch._dynamo.mark_dynamic(x, 1)                        # moondream.py:303
            T, pos = x.size(1), enc.pos
            hidden = self._prefill(x, self.attn_mask[:, :, pos:pos + T, :],
                                   torch.arange(pos, pos + T, dtype=torch.long, device=self.device))
            nxt = torch.argmax(self.lm_head(hidden), dim=-1).unsqueeze(1)
            pos += T
            mask = torch.zeros(1, 1, self.cfg.text.max_context, device=self.device, dtype=torch.bool)
            mask[:, :, :pos] = 1
            pos_ids = torch.tensor([pos], device=self.device, dtype=torch.long)
            n = 0
            while (tok := nxt.item()) != tk.eos_id and n < max_tokens:
                out.append(tok)
                emb = self.embed(nxt)
                mask[:, :, pos], pos_ids[0] = 1, pos
                logits, _ = self._decode_one_tok(emb, mask, pos_ids)
                logits = logits.clone()
                logits[:, tk.answer_id] = float("-inf")
                pos += 1
                nxt = torch.argmax(logits, dim=-1).unsqueeze(1)
                n += 1
        return out


def _sync():
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def measure(orc: SeamOracle, cfg, n_img: int, n_tok: int, warmup: int):
    from moondream_b200 import synth

    def one(i):
        img = synth.synthetic_image(i, 378, 378)
        prompt = synth.synthetic_prompt(i, 32, cfg.text.vocab_size)
        t0 = time.perf_counter()
        enc = orc.encode(img)
        _sync()
        t1 = time.perf_counter()
        toks = orc.answer(enc, prompt, n_tok)
        _sync()
        return t1 - t0, time.perf_counter() - t1, toks

    for i in range(warmup):
        one(i)
    _sync()
    t0 = time.perf_counter()
    runs = [one(i) for i in range(n_img)]
    _sync()
    dt = time.perf_counter() - t0
    return {"images": n_img, "tokens_per_image": n_tok, "seconds": dt, "images_per_s": n_img / dt,
            "encode_ms_mean": 1e3 * sum(r[0] for r in runs) / n_img,
            "decode_tokens_per_s": sum(len(r[2]) for r in runs) / max(1e-9, sum(r[1] for r in runs)),
            "first_tokens": runs[0][2][:8]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=8)
    ap.add_argument("--tokens", type=int, default=64)
    ap.add_argument("--model", default="moondream-2b")
    ap.add_argument("--head-peak", type=float, default=0.0)
    ap.add_argument("--compile", action="store_true")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--backend", default="inductor")
    args = ap.parse_args()
    from moondream_b200 import config as C, synth

    cfg = C.preset(args.model)
    sd = synth.synthetic_state_dict(cfg, 0, head_peak=args.head_peak)
    orc = SeamOracle(cfg, sd, device=args.device)
    res = {"what": "oracle port of the reference (bit-identical arithmetic), torch ops on " + args.device +
                   ", batch-1 sequential, one .item() per token (moondream.py:482)",
           "gpu": torch.cuda.get_device_name(0) if args.device.startswith("cuda") else "cpu",
           "torch": torch.__version__, "measured_in_run": True}
    if not args.compile:
        res["mode"] = "eager"
        res.update(measure(orc, cfg, args.images, args.tokens, warmup=1))
    else:
        res["mode"] = "torch.compile (the reference's compile(): _vis_enc/_prefill fullgraph, _decode_one_tok reduce-overhead)"
        t0 = time.perf_counter()
        torch._dynamo.reset()
        orc.compile(args.backend)
        res.update(measure(orc, cfg, args.images, args.tokens, warmup=3))
        res["compile_and_warmup_s"] = time.perf_counter() - t0 - res["seconds"]
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
