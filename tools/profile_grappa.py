#!/usr/bin/env python
"""Minimal driver for ncu: GRAPPA weight application at the prostate file shape (30 x 16 x 640 x 451, R = 2 + 24 ACS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mri_acl_imagesegmentation_adsp_b200.adapters import recon_cabi
from mri_acl_imagesegmentation_adsp_b200.prostate.grappa import Grappa
S, C, RO, PE = (int(a) for a in (sys.argv[1:5] if len(sys.argv) >= 5 else (30, 16, 640, 451)))
dev = torch.device("cuda")
keep = np.zeros(PE, dtype=bool); keep[::2] = True; keep[(PE - 24) // 2:(PE - 24) // 2 + 24] = True
g = torch.Generator(device=dev).manual_seed(78)
k = torch.view_as_complex(torch.randn((S, C, RO, PE, 2), device=dev, generator=g))
k[..., torch.from_numpy(~keep).to(dev)] = 0
gr = Grappa(np.transpose(k[0].cpu().numpy(), (2, 0, 1)), (5, 5), 1)
kv = gr.kernel_var_dict
rng = np.random.default_rng(5)
wd = {int(i): (0.05 * (rng.standard_normal((C, int(kv["patches"][i].sum()))) + 1j * rng.standard_normal((C, int(kv["patches"][i].sum()))))).astype(np.complex64)
      for i in kv["patch_indices"]}
plan = gr._device_plan(dev, lanes_along_x=os.environ.get("LANES_X", "1") == "1")
W = gr._pack_weights([wd] * S, plan)
lib = recon_cabi.library()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(4):
    if it == 1: ev[0].record()
    lib.grappa_apply(k.data_ptr(), C * RO * PE, 1, PE, RO * PE, S, PE, RO, C, 5, 5, plan["hole_xy"].data_ptr(), plan["n_items"],
                     plan["item_geom"].data_ptr(), plan["item_first"].data_ptr(), plan["item_count"].data_ptr(), plan["src_start"].data_ptr(),
                     plan["src_off"].data_ptr(), plan["max_src"], plan["w_start"].data_ptr(), W.data_ptr(), W.shape[1], 0)
ev[1].record()
torch.cuda.synchronize()
flops = sum(8.0 * C * int(kv["patches"][g_][..., 0].sum()) * C * len(kv["holes_x"][g_]) for g_ in plan["geoms"]) * S
ms = ev[0].elapsed_time(ev[1]) / 3
print(f"grappa apply: {ms:.3f} ms, {flops / ms / 1e9:.1f} TFLOP/s, items {plan['n_items']}, geoms {len(plan['geoms'])}, max_src {plan['max_src']}")
