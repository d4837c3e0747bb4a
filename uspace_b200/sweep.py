"""Mirror of the write branch of ``sample_for_hspace_vis`` (tools/utils_vis.py:180-201): one latent batch decoded
once per ``write_scale`` and laid out "(b s)" - here as ONE batched integration of B * len(write_scales) samples when
the solver is the fixed grid (samples are independent, so each row equals a separate ``decode`` bit for bit; the
larger batch fills the last wave of GEMM tiles), and as the reference's loop otherwise (an adaptive solver's step
size depends on the whole batch)."""
from __future__ import annotations

import torch

from .engine import time_grid
from .flow_matching import CNF, build_delta_table


def sample_write_scales(cnf, input_z: torch.Tensor, write_scales, cond=None, **kwargs) -> torch.Tensor:
    """Latents [B * S, C, H, W] in "(b s)" order; ``kwargs`` are the reference's dissection kwargs
    (dissect_task, dissect_name in {"write_attr", "write_pca"}, write_path_root, ith_attr, t_edit, edit_loc,
    solver_kwargs) - what ``sample_fn`` forwards to ``score_model.decode`` (dissect_lfm.py:130-141)."""
    if kwargs.get("dissect_name") not in ("write_attr", "write_pca"):
        raise NotImplementedError(f"dissect_name should be write_attr or write_pca, but got: {kwargs.get('dissect_name')}")
    kwargs.pop("write_scale", None)
    write_scales = [float(s) for s in write_scales]
    sk = kwargs["solver_kwargs"]
    cond_kw = dict(y=cond) if isinstance(cnf, CNF) else dict(context=cond)
    if sk["solver"] == "fixed":
        ode = cnf.get_ode_kwargs(**kwargs)
        h = ode["options"]["step_size"]
        table, loc = build_delta_table(time_grid(0.0, 1.0, h), input_z.shape[1:], **kwargs)
        if table is not None:
            net = cnf.net.module if hasattr(cnf.net, "module") else cnf.net
            out = net.engine().sample_sweep(input_z, write_scales, 0.0, 1.0, h, ode["method"], delta_table=table,
                                            t_edit=float("inf"), edit_loc=loc, **cond_kw)
            return out.reshape(-1, *input_z.shape[1:])
    outs = [cnf.decode(input_z, cond, write_scale=s, **kwargs) for s in write_scales]
    return torch.stack(outs, dim=1).reshape(-1, *input_z.shape[1:])
