"""Run under torchrun on N GPUs: `ShardedModel` (the multi-GPU product entry point) must return, on every rank, exactly
what a single GPU returns for the same request — captions (greedy), detect boxes — with images of mixed sizes so the
crop-count sharding is ragged.  Prints one line per rank."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from moondream_b200 import config as C, synth  # noqa: E402
from moondream_b200.moondream import MoondreamModel  # noqa: E402
from moondream_b200.parallel import ShardedModel  # noqa: E402
from oracle.reference_shim import StubTokenizer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = C.tiny()
sd = synth.synthetic_state_dict(cfg, 0)
model = MoondreamModel(cfg, tokenizer=StubTokenizer(cfg.text.vocab_size), max_batch=8, device=f"cuda:{local}")
model.load_state_dict(sd)
sizes = [(378, 378), (800, 600), (500, 700), (378, 378), (1000, 1200), (300, 200), (756, 756)]
images = [synth.synthetic_image(i, h, w) for i, (h, w) in enumerate(sizes)]
settings = {"temperature": 0, "max_tokens": 12}
sm = ShardedModel(model)
parts = sm.sharded.plan(images)
got = sm.caption_batch(images, "short", settings=settings)
want = model.caption_batch(images, "short", settings=settings)
assert got == want, (rank, got, want)
qs = [f"{11 + i} 12" for i in range(len(images))]
assert sm.query_batch(images, qs, settings=settings) == model.query_batch(images, qs, settings=settings)
det = sm.detect_batch(images, ["17 23"] * len(images), settings={"max_objects": 3})
ref = model.detect_batch(images, ["17 23"] * len(images), settings={"max_objects": 3})
assert len(det) == len(ref)
for a, b in zip(det, ref):
    assert len(a["objects"]) == len(b["objects"])
    for x, y in zip(a["objects"], b["objects"]):
        assert all(abs(x[k] - y[k]) < 1e-6 for k in x), (x, y)
pts = sm.point_batch(images, ["17 23"] * len(images), settings={"max_objects": 2})
assert [len(p["points"]) for p in pts] == [len(p["points"]) for p in model.point_batch(images, ["17 23"] * len(images), settings={"max_objects": 2})]
dist.barrier()
print(f"rank {rank}/{world}: sharded == local for {len(images)} images; plan {parts}", flush=True)
dist.destroy_process_group()
