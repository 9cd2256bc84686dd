"""The portable oracle against the fixtures generated from the REAL reference
(oracle/make_golden.py, tests/golden/*.pt).  CPU only; sized to run in about a minute."""
import pytest
import torch

from oracle.golden_cases import CASES, case_batch, load_golden
from util import run_oracle

CPU_CASES = [c for c in CASES if c != 'dense_1600'] + ['dense_1600']


def _check_digest(t, dg, rtol):
    flat = t.reshape(-1)
    assert tuple(t.shape) == tuple(dg['shape'])
    assert torch.allclose(flat[dg['idx']], dg['val'], rtol=rtol, atol=0)
    assert abs(float(flat.double().sum()) - dg['sum']) <= rtol * max(dg['abssum'], 1e-30)
    assert abs(float((flat.double() ** 2).sum()) - dg['sqsum']) <= 10 * rtol * max(dg['sqsum'], 1e-30)
    assert int((flat != 0).sum()) == dg['nnz']
    if 'nz_idx' in dg:
        assert torch.equal((flat != 0).nonzero().squeeze(1), dg['nz_idx'])
        assert torch.allclose(flat[dg['nz_idx']], dg['nz_val'], rtol=rtol, atol=0)


@pytest.mark.parametrize('name', CPU_CASES)
def test_oracle_reproduces_reference_fixture(name):
    gold = load_golden(name)
    if gold['torch_version'] != torch.__version__:
        pytest.skip('fixture generated with another torch build')
    batch = case_batch(name)
    o = run_oracle(batch)
    for i in range(batch.num_imgs):
        assert torch.equal(o['cls_inds'][i], gold['cls_inds'][i])
        assert torch.equal(o['box_inds'][i], gold['box_inds'][i])
        assert torch.equal(o['keep'][i], gold['keep'][i])
        gi = o['gt_inds'][i]
        assert torch.equal((gi > 0).nonzero().squeeze(1), gold['pos'][i])
        assert torch.equal(gi[gi > 0], gold['pos_gt'][i])
        assert int((gi < 0).sum()) == gold['num_invalid'][i]
    for k, v in gold['losses'].items():
        assert o['losses'][k] == v, k              # same ops on the same CPU: bit-identical
    for l in range(5):
        _check_digest(o['g_cls'][l], gold['g_cls'][l], 0.0)
        _check_digest(o['g_box'][l], gold['g_box'][l], 0.0)


def test_tiny_fixture_inputs_are_reproducible():
    """The committed inputs of the tiny case equal what the seeded generator rebuilds."""
    gold = load_golden('tiny_40_40')
    b = case_batch('tiny_40_40')
    for key in ('t_cls', 't_box', 's_cls', 's_box', 'gt_bboxes', 'gt_labels'):
        for x, y in zip(getattr(b, key), gold['inputs'][key]):
            assert torch.equal(x, y), key
    o = run_oracle(b)
    for l in range(5):
        assert torch.equal(o['g_cls'][l], gold['g_cls_full'][l])
        assert torch.equal(o['g_box'][l], gold['g_box_full'][l])
