"""TEST INFRASTRUCTURE ONLY -- golden vectors for SimOTA's dynamic-k matching, written by the
UNMODIFIED reference method ``SimOTABEVAssigner.dynamic_k_matching``
(``core/bbox/assigners/sim_ota_3d_assigner.py:184-211``, loaded by ``oracle/ref_loader.py`` under
stub imports).  Inputs are tie-free (distinct random values), so ``torch.topk``'s unspecified tie
order cannot matter.  Build container only:

    python oracle/make_simota_golden.py        ->  tests/golden/gd_simota_golden.npz
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402


def main():
    cls = ref_loader.load_reference_simota()
    g = torch.Generator().manual_seed(0)
    out, manifest = {}, []
    for cid, (n, m, topk, scale) in enumerate([(200, 7, 10, 1.0), (64, 12, 10, 0.5), (9, 3, 10, 1.0),
                                               (500, 31, 5, 2.0), (300, 1, 10, 1.0),
                                               (1000, 40, 13, 0.3), (50, 50, 10, 1.5)]):
        # similarity in (0, 1): a few boxes close to each GT, most far away, all distinct
        d = torch.rand(n, m, generator=g, dtype=torch.float64) * scale
        near = torch.rand(n, m, generator=g, dtype=torch.float64) < 0.08
        d = torch.where(near, d * 0.05, 0.3 + d)
        ious = 1.0 / (1.0 + d)                       # tau = 1 similarity of a distance d
        cost = 1.0 - ious                            # a cost monotone in the distance
        self = cls(candidate_topk=topk)
        valid = torch.ones(n, dtype=torch.bool)
        matched_ious, matched_gt = self.dynamic_k_matching(cost.clone(), ious.clone(), m, valid)
        assigned = torch.zeros(n, dtype=torch.int64)
        assigned[valid] = matched_gt + 1             # sim:112 (valid was updated in place, sim:206)
        full = torch.zeros(n, dtype=torch.float64)
        full[valid] = matched_ious
        out[f'{cid}/cost'] = cost.numpy()
        out[f'{cid}/ious'] = ious.numpy()
        out[f'{cid}/assigned'] = assigned.numpy()
        out[f'{cid}/matched_ious'] = full.numpy()
        manifest.append(dict(id=cid, n=n, m=m, candidate_topk=topk))
    out['manifest'] = np.frombuffer(json.dumps(manifest).encode(), dtype=np.uint8)
    path = os.path.join(ROOT, 'tests', 'golden', 'gd_simota_golden.npz')
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), 'bytes,', len(manifest), 'cases')


if __name__ == '__main__':
    main()
