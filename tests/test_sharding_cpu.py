"""World-size-2 test of the image-wise sharding host logic on CPU (gloo): the N>1 path of bench.py / denoise_stream
has no data-path collective, only round-robin ownership and a host-side gather."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import importlib
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sharding = importlib.import_module("ai-cv-automation-elect-micr_b200.sharding")
    rng = np.random.default_rng(7)
    images = [rng.random((6, 5)) for _ in range(7)]
    calls = []

    def fake_denoise(img):   # stands in for Denoiser.denoise: deterministic per image, records who ran it
        calls.append(float(img.sum()))
        return 1.0 - img

    out = sharding.run_sharded(images, fake_denoise, rank, world)
    ok = all(np.array_equal(o, 1.0 - i) for o, i in zip(out, images))

    class FakeDenoiser:      # the stream form: a rank's same-sized images go through ONE denoise_many call (emd_denoise_stream)
        batches = []

        def denoise(self, img, **kw):
            raise AssertionError("same-sized images must take the batched route")

        def denoise_many(self, imgs, **kw):
            self.batches.append(len(imgs))
            return [1.0 - i * kw.get("overlap", 1) for i in imgs]

    fd = FakeDenoiser()
    out2 = sharding.denoise_stream(fd, images, rank, world, overlap=2)
    ok = ok and all(np.array_equal(o, 1.0 - 2 * i) for o, i in zip(out2, images)) and fd.batches == [len(sharding.shard_indices(7, rank, world))]
    ragged = images[:2] + [rng.random((4, 4)), rng.random((4, 4))]   # each rank owns one 6x5 and one 4x4: one denoise call per image
    seen = []

    class PerImage:
        def denoise(self, img, **kw):
            seen.append(img.shape)
            return img + 1.0

        def denoise_many(self, imgs, **kw):
            raise AssertionError("mixed sizes cannot be batched")

    out3 = sharding.denoise_stream(PerImage(), ragged, rank, world)
    ok = ok and all(np.array_equal(o, i + 1.0) for o, i in zip(out3, ragged))
    q.put((rank, ok, len(calls), sharding.shard_indices(len(images), rank, world)))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_ownership():
    sys.path.insert(0, ROOT)
    import importlib
    sharding = importlib.import_module("ai-cv-automation-elect-micr_b200.sharding")
    for n, w in [(0, 1), (7, 2), (100, 8), (3, 8)]:
        owned = [sharding.shard_indices(n, r, w) for r in range(w)]
        assert sorted(k for o in owned for k in o) == list(range(n))
        assert all(k % w == r for r, o in enumerate(owned) for k in o)
    with pytest.raises(ValueError):
        sharding.shard_indices(4, 2, 2)


def test_two_rank_gather_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] and res[1][1]                       # both ranks hold every image's result, in item order
    assert (res[0][2], res[1][2]) == (4, 3)              # 7 images: rank 0 ran 4, rank 1 ran 3 -- nothing ran twice
    assert res[0][3] == [0, 2, 4, 6] and res[1][3] == [1, 3, 5]
