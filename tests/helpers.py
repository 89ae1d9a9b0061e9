import hashlib
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def checksum(tensors) -> str:
    h = hashlib.sha256()
    for t in tensors:
        h.update(np.ascontiguousarray(t.detach().cpu().numpy()).tobytes())
    return h.hexdigest()


def batch_checksum(b) -> str:
    return checksum([b[k] for k in sorted(b) if torch.is_tensor(b[k])]
                    + [torch.from_numpy(np.asarray(m)) for m in b["rot_node_mask"]])


def conditioning(b, t=0.7, tr_sigma=1.5, rot_norm=0.8, tor_norm2=0.5):
    B = b["num_graphs"]
    return dict(t=torch.full((B,), t), tr_sigma=torch.full((B,), tr_sigma), rot_score_norm=torch.full((B, 1), rot_norm),
                tor_score_norm2=torch.full((int(b["tor_edge_mask"].sum()),), tor_norm2),
                sc_tor_score_norm2=torch.full(tuple(b["sc_torsion_edge_mask"].shape), tor_norm2) * b["sc_torsion_edge_mask"])


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def rmsd(a, b):
    return float(torch.sqrt(((a.double() - b.double()) ** 2).sum(-1).mean()))
