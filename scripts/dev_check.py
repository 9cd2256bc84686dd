"""Developer diagnostic: CUDA path vs oracle on one synthetic batch, printed verbosely.
Usage: python scripts/dev_check.py [case kwargs as python dict literal]"""
import ast
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from erd_b200.synth import make_batch  # noqa: E402
from util import run_cuda, run_oracle, rel_err  # noqa: E402


def main():
    kw = dict(num_imgs=2, img_hw=(800, 1333), ori=40, seed=1234)
    if len(sys.argv) > 1:
        kw.update(eval(sys.argv[1]))
    print('case', kw)
    b = make_batch(**kw)
    t = time.time()
    o = run_oracle(b)
    print(f'oracle {time.time() - t:.2f}s')
    t = time.time()
    c = run_cuda(b)
    print(f'cuda {time.time() - t:.2f}s  avg cuda {c["avg"]} oracle {o["avg"]}')
    print('thr', c['thr'].tolist(), 'oracle', o['report']['cls_thr'], o['report']['box_thr'])
    for i in range(b.num_imgs):
        for key in ('cls_inds', 'box_inds', 'keep'):
            same = torch.equal(c[key][i], o[key][i])
            print(f'img{i} {key}: cuda {len(c[key][i])} oracle {len(o[key][i])} equal={same}')
            if not same and key != 'keep':
                sc, so = set(c[key][i].tolist()), set(o[key][i].tolist())
                print('   only cuda', sorted(sc - so)[:10], 'only oracle', sorted(so - sc)[:10])
        gc, go = c['gt_inds'][i], o['gt_inds'][i]
        print(f'img{i} gt_inds equal={torch.equal(gc, go)} pos cuda {(gc > 0).sum().item()} oracle {(go > 0).sum().item()} '
              f'invalid {(gc < 0).sum().item()}/{(go < 0).sum().item()}')
        if not torch.equal(gc, go):
            d = (gc != go).nonzero().flatten()[:10]
            print('   diff at', d.tolist(), gc[d].tolist(), go[d].tolist())
    for k in o['losses']:
        for x, y in zip(c['losses'][k], o['losses'][k]):
            print(f'{k}: cuda {x:.8f} oracle {y:.8f} rel {abs(x - y) / max(abs(y), 1e-12):.2e}')
    for l in range(5):
        print(f'level{l} g_cls rel {rel_err(c["g_cls"][l], o["g_cls"][l]):.2e} '
              f'g_box rel {rel_err(c["g_box"][l], o["g_box"][l]):.2e} '
              f'|g_cls| {float(o["g_cls"][l].abs().max()):.3e} |g_box| {float(o["g_box"][l].abs().max()):.3e}')


if __name__ == '__main__':
    main()
