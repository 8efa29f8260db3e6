#!/usr/bin/env python
"""Generates tests/golden/cornell_oracle_golden.npz from the CPU oracle.

The reference has no golden vectors for this path, and it cannot be run in
this image, so these are regression pins of the ORACLE (first-hit ids, t bits, traversal
counters and a small radiance image on the cornell-box fixture).  Re-run after a deliberate
spec change:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from loupiote_b200 import _ffi, scenes  # noqa: E402
from oracle import oracle as O  # noqa: E402

V_FOV = 0.78539816339


def main():
    c = scenes.cornell_box()
    osc = O.OracleScene(c["scene"])
    cam = O.camera_from_view(c["view"], 64, 64, V_FOV)
    inst, prim, t, _, st = O.first_hit_image(osc, cam, 1)
    cam2 = O.camera_from_view(c["view"], 32, 32, V_FOV)
    cfg = _ffi.RenderConfig()
    _ffi.lib().lp_render_config_default(cfg)
    cfg.max_bounces, cfg.jitter, cfg.seed = 4, 1, 1
    acc, rs = O.render(osc, cam2, cfg, 4)
    np.savez_compressed(
        Path(__file__).resolve().parent / "cornell_oracle_golden.npz",
        inst=inst, prim=prim, t_bits=t.view(np.uint32),
        stats=np.array([st["n_int"], st["n_tri"], st["n_inst"]], dtype=np.int64),
        radiance_32x32_4spp_4b=(acc[..., :3] / acc[..., 3:4]).astype(np.float32),
        ray_counts=np.array([rs["primary"], rs["bounce"], rs["shadow"]], dtype=np.int64))
    print("wrote golden:", st, rs["primary"], rs["bounce"], rs["shadow"])

    # textured scene under the procedural probe
    c = scenes.textured_scene()
    osc = O.OracleScene(c["scene"], env_color=c["env_color"], probe=c["probe"])
    cam3 = O.camera_from_view(c["view"], 64, 36, V_FOV)
    acc, rs, gb, _ = O.render(osc, cam3, cfg, 4, want_gbuffer=True)
    np.savez_compressed(
        Path(__file__).resolve().parent / "textured_oracle_golden.npz",
        radiance_64x36_4spp_4b=(acc[..., :3] / acc[..., 3:4]).astype(np.float32),
        ray_counts=np.array([rs["primary"], rs["bounce"], rs["shadow"]], dtype=np.int64),
        gbuffer_albedo=gb[..., 3])
    print("wrote textured golden:", rs["primary"], rs["bounce"], rs["shadow"])


if __name__ == "__main__":
    main()
