"""The offline step between the hook's "read" and "write_attr" modes: semantic directions from attribute labels.

Mirror of ``extract_hspace_feat_unet_by_attr`` / ``cal_delta_direction`` / ``cal_latentz_delta``
(tools/utils_attr.py:123-206) on the same on-disk formats:

  * ``{batch_id}_{t:.2f}.npy``   [b, C, W, H]  one file per batch and evaluation time, written by the "read" mode
                                  (libs/dissection.py:126-136; here: ``CNF.encode(..., dissect_name="read")``)
  * ``latents.npy.npz``          ``latent`` [B, C, W, H], ``attr`` [B, 40 | 11] in {0, 1} (dissect_lfm.py:224-228)
  * ``delta_{t:.2f}.npy``        [attr_dim, C, W, H] = mean(feat | attr == 1) - mean(feat | attr == 0), read back by
                                  the "write_attr" mode (libs/dissection.py:143-150)
  * ``delta_latentz.npy``        the same statistic on the encoded latents (the "write_x0" direction)

It is host-side numpy like the reference (file-format glue, IO-bound); one timestep is resident at a time instead
of the reference's [B, T, C, W, H] stack."""
from __future__ import annotations

import os

import numpy as np

_ATTR_DIMS = (40, 11)   # CelebA_ATTR40 / FFHQ_ATTR11 (tools/utils_attr.py:14-91): anything else raises there too


def should_ignore(name: str) -> bool:
    """tools/utils_attr.py:94-101."""
    return name.startswith("pca") or name.startswith("latent") or name.startswith("delta")


def save_latents(read_path_root: str, latent, attr) -> str:
    """dissect_lfm.py:224-228: np.savez(<root>/latents.npy, latent=, attr=) (numpy appends ".npz")."""
    os.makedirs(read_path_root, exist_ok=True)
    np.savez(os.path.join(read_path_root, "latents.npy"), latent=np.asarray(latent), attr=np.asarray(attr))
    return os.path.join(read_path_root, "latents.npy.npz")


def cal_delta_direction(attr_id: int, attrs: np.ndarray, feats: np.ndarray) -> np.ndarray:
    """[1, ...] = mean over samples with attrs[:, attr_id] == 1 minus mean over those with == 0."""
    if attrs.shape[1] not in _ATTR_DIMS:
        raise ValueError("unknown attr dim", len(attrs))
    a = attrs[:, attr_id]
    pos = np.mean(feats[a == 1], axis=0, keepdims=True)
    neg = np.mean(feats[a == 0], axis=0, keepdims=True)
    return pos - neg


def cal_latentz_delta(read_path_root: str, latent_file: str = "latents.npy.npz") -> str:
    data = np.load(os.path.join(read_path_root, latent_file))
    attrs, latent = data["attr"], data["latent"]
    out = np.concatenate([cal_delta_direction(i, attrs, latent) for i in range(attrs.shape[1])], axis=0)
    path = os.path.join(read_path_root, "delta_latentz")
    np.save(path, out)
    return path + ".npy"


def extract_deltas_by_attr(read_path_root: str, batch_num: int, latent_file: str = "latents.npy.npz",
                           cal_latentz_delta_only: bool = False):
    """Writes delta_{t}.npy for every evaluation time found under ``read_path_root``; returns the sorted time strings."""
    attrs = np.load(os.path.join(read_path_root, latent_file))["attr"]
    if cal_latentz_delta_only:
        cal_latentz_delta(read_path_root, latent_file)
        return []
    names = [n for n in os.listdir(read_path_root) if not should_ignore(n)]
    timesteps = sorted({n.split("_")[1].replace(".npy", "") for n in names})
    for ts in timesteps:
        feats = [np.load(os.path.join(read_path_root, f"{b}_{ts}.npy")) for b in range(batch_num)]
        if not feats:
            raise ValueError("**** empty feat", ts)
        feat = np.concatenate(feats, axis=0)                       # [B, C, W, H]
        delta = np.concatenate([cal_delta_direction(i, attrs, feat) for i in range(attrs.shape[1])], axis=0)
        np.save(os.path.join(read_path_root, f"delta_{ts}"), delta)   # [attr_dim, C, W, H]
    return timesteps
