#!/usr/bin/env python
"""Denoise every 32-bit float TIFF micrograph of a directory (BASELINE config 4 from disk).
  python tools/denoise_tiffs.py --in micrographs/ --out denoised/ [--checkpoint <TF checkpoint dir | weights.emdw>] [--overlap 80]
Under `python -m torch.distributed.run --nproc-per-node N` each rank takes the files k with k % N == rank on its own GPU
(no collective: the path shards by micrograph)."""
import argparse
import glob
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--in", dest="src", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--variant", default="A", choices=["A", "B"])
    ap.add_argument("--overlap", type=int, default=80)
    ap.add_argument("--mode", default="fp16", choices=["bf16", "fp16", "fp32"])
    a = ap.parse_args()
    import denoiser
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    paths = sorted(glob.glob(os.path.join(a.src, "*.tif")) + glob.glob(os.path.join(a.src, "*.tiff")))
    d = denoiser.Denoiser(checkpoint_loc=a.checkpoint, device=int(os.environ.get("LOCAL_RANK", 0)), mode=a.mode, variant=a.variant)
    t0 = time.perf_counter()
    done = denoiser.emd.micrograph_io.denoise_files(d, paths, a.out, rank=rank, world=world, overlap=a.overlap)
    dt = time.perf_counter() - t0
    print(f"rank {rank}/{world}: {len(done)} of {len(paths)} micrographs in {dt:.2f} s")


if __name__ == "__main__":
    main()
