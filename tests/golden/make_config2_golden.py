#!/usr/bin/env python
"""Generates tests/golden/config2_oracle_ref.npz: the CPU oracle's converged cornell box that
the config-2 image gate of tests/test_gpu_a_bench_size.py compares the GPU render with
(BASELINE config 2; SURVEY 8(d) "Converged image").

The reference holds no golden image and cannot run here, so this is the
ORACLE's image: 1024 spp, 8 bounces, 240x135, seed 1001 -- plus the oracle's own 256-spp image
with another seed (2002) scored against it, which calibrates the gate:
    RMSE(gpu_256, ref) <= 1.25 RMSE(oracle_256, ref) + 0.002,  |mean(gpu) - mean(ref)| <= 1 %.
About 40 s on 8 cores:  python tests/golden/make_config2_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from loupiote_b200 import _ffi, metrics, scenes  # noqa: E402
from oracle import oracle as O  # noqa: E402

V_FOV = 0.78539816339
W, H, BOUNCES = 240, 135, 8


def main():
    c = scenes.cornell_box()
    osc = O.OracleScene(c["scene"])
    cam = O.camera_from_view(c["view"], W, H, V_FOV)
    cfg = _ffi.RenderConfig()
    _ffi.lib().lp_render_config_default(cfg)
    cfg.max_bounces, cfg.jitter = BOUNCES, 1
    cfg.seed = 1001
    acc, _ = O.render(osc, cam, cfg, 1024)
    ref = (acc[..., :3] / acc[..., 3:4]).astype(np.float32)
    cfg.seed = 2002
    acc, _ = O.render(osc, cam, cfg, 256)
    o256 = (acc[..., :3] / acc[..., 3:4]).astype(np.float32)
    rmse = metrics.normalised_rmse(o256, ref)
    flip = metrics.mean_flip(ref, o256)
    np.savez_compressed(Path(__file__).resolve().parent / "config2_oracle_ref.npz", ref=ref,
                        rmse_oracle_256=np.float64(rmse), flip_oracle_256=np.float64(flip),
                        meta=np.array([W, H, BOUNCES, 1024, 1001, 256, 2002], dtype=np.int64))
    print(f"wrote config2_oracle_ref.npz: rmse(oracle 256) {rmse:.5f}, mean FLIP {flip:.5f}, "
          f"mean radiance {ref.mean():.5f}")


if __name__ == "__main__":
    main()
