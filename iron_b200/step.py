"""One stage-2 step of the hot path, as the reference's training loop composes it (render_surface.py:533-653,
restricted to the rows in scope): trace -> shade (is_training) -> L2 image loss + eikonal -> backward.
This is the unit bench.py times and tests/test_step.py compares against the oracle's stage2_step."""
from __future__ import annotations

import torch

from .raytracer import render_camera


def stage2_step(sdf_network, color_network_dict, raytracer, render_fn, camera, target, eik_points, eik_weight=0.1,
                max_num_rays=50000, fill_holes=False, handle_edges=False, dense_shading=False):
    """Leaves gradients in .grad of every parameter; returns (loss, results).  fill_holes / handle_edges = True is the
    reference drivers' default configuration (render_surface.py:521-549)."""
    results = render_camera(camera, sdf_network, raytracer, color_network_dict, render_fn, fill_holes=fill_holes,
                            handle_edges=handle_edges, is_training=True, dense_shading=dense_shading)
    mask = results["convergent_mask"]
    if handle_edges:
        mask = mask | results["edge_mask"]                                # render_surface.py:566-567
    eg = sdf_network.gradient(eik_points).view(-1, 3)                 # render_surface.py:580-583
    eik_cnt = eg.shape[0]
    eik = ((eg.norm(dim=-1) - 1) ** 2).sum()
    img = ((results["color"] - target) ** 2).sum() / float(mask.numel())
    hn = results["normal"].reshape(-1, 3)
    hm = mask.reshape(-1, 1).float()
    n_hit = mask.sum()
    eik = eik + (((hn.norm(dim=-1, keepdim=True) - 1) ** 2) * hm).sum()   # hit normals (:601-603), no host sync
    if "edge_pos_neg_normal" in results:                                  # :604-607
        en = results["edge_pos_neg_normal"]
        eik_cnt += en.shape[0]
        eik = eik + ((en.norm(dim=-1) - 1) ** 2).sum()
    loss = img + eik / (eik_cnt + n_hit) * eik_weight
    loss.backward()
    return loss.detach(), results
