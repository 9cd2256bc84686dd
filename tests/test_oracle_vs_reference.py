"""Live check of the oracle port against the reference's own code loaded by path.  Runs only
where the reference tree exists (the build container); skipped on the GPU box."""
import pytest
import torch

from oracle import ref_by_path

pytestmark = pytest.mark.skipif(not ref_by_path.available(), reason='reference tree not present')


@pytest.mark.parametrize('kw', [
    dict(num_imgs=2, img_hw=(256, 320), ori=40, seed=5, gt_size_pow=2.0),
    dict(num_imgs=2, img_hw=(256, 320), ori=70, seed=6, mode='trained', gt_size_pow=2.0),
    dict(num_imgs=2, img_hw=(320, 320), ori=40, seed=7, num_gt=[0, 6], mode='trained',
         pad_shapes=[(320, 320), (288, 200)], gt_size_pow=2.5),
])
def test_port_is_bit_identical_to_reference(kw):
    from erd_b200.synth import make_batch
    from oracle.make_golden import run_reference
    from util import run_oracle
    b = make_batch(**kw)
    r, o = run_reference(b), run_oracle(b)
    for i in range(b.num_imgs):
        assert torch.equal(r['cls_inds'][i], o['cls_inds'][i]) and torch.equal(r['box_inds'][i], o['box_inds'][i])
        assert torch.equal(r['keep'][i], o['keep'][i]) and torch.equal(r['gt_inds'][i], o['gt_inds'][i])
    assert r['losses'] == o['losses']
    for l in range(5):
        assert torch.equal(r['g_cls'][l], o['g_cls'][l]) and torch.equal(r['g_box'][l], o['g_box'][l])


def test_reference_own_goldens_under_the_shim():
    """The reference's classes, executed by path, reproduce their own test goldens."""
    ref = ref_by_path.load_reference()
    a = ref.ATSSAssigner(topk=9)
    priors = torch.FloatTensor([[0, 0, 10, 10], [10, 10, 20, 20], [5, 5, 15, 15], [32, 32, 38, 42]])
    gt = ref.InstanceData(bboxes=torch.FloatTensor([[0, 0, 10, 9], [0, 10, 10, 19]]), labels=torch.LongTensor([2, 3]))
    res = a.assign(ref.InstanceData(priors=priors), [4], gt)
    assert res.gt_inds.tolist() == [1, 0, 0, 0]
    b1 = torch.FloatTensor([[0, 0, 10, 10], [10, 10, 20, 20], [32, 32, 38, 42]])
    b2 = torch.FloatTensor([[0, 0, 10, 20], [0, 10, 10, 19], [10, 10, 20, 20]])
    g = ref.bbox_overlaps(b1, b2, 'giou', is_aligned=True, eps=1e-7)
    assert torch.allclose(g, torch.tensor([0.5, -0.05, -0.8214]), atol=1e-4)
