"""GPU tests of the bulk code extraction path: CodeExtractor must return exactly the indices
`get_code_indices` returns batch by batch (vq_codes_lmdb.py:58-96 semantics) and the across-batch codebook usage the
reference computes with a host-side cat + torch.unique (main.py:261-330)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mv():
    import movae_b200
    return movae_b200


@pytest.mark.parametrize("dtype", [None, torch.int16, torch.int32, torch.int64])
def test_extractor_matches_per_batch_indices_and_unique(mv, dtype):
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(5)
    vq = mv.VectorQuantizer(512, 64).to(dev)
    with torch.no_grad():
        vq.embedding.weight.copy_(0.5 * torch.randn(512, 64, generator=g, device=dev))
    ex = mv.CodeExtractor(vq, code_dtype=dtype, ring=2)
    batches = [0.5 * torch.randn(B, 64, H, W, generator=g, device=dev) for B, H, W in ((7, 8, 8), (3, 16, 16), (1, 5, 7), (7, 8, 8), (2, 3, 3))]
    expect = []
    for z in batches:
        ex.push(z)
        expect.append(vq.get_code_indices(z))
    per = ex.finish(per_batch=True)
    flat = ex.finish()
    assert flat.dtype == (dtype or torch.int16) and not flat.is_cuda
    assert len(per) == len(batches)
    for z, got, want in zip(batches, per, expect):
        assert tuple(got.shape) == (z.shape[0], z.shape[2], z.shape[3])          # [B, h, w] like vq_vae.py:419-421
        assert torch.equal(got.reshape(-1).to(torch.int64), want.cpu())
    assert torch.equal(flat.to(torch.int64), torch.cat(expect).cpu())
    n_unique = torch.unique(torch.cat(expect)).numel()                           # the reference's host-side accounting
    assert int(ex.usage_count().item()) == n_unique
    assert ex.usage_percentage() == pytest.approx(100.0 * n_unique / 512)


def test_extractor_rejects_bad_arguments(mv):
    vq = mv.VectorQuantizer(512, 64).cuda()
    with pytest.raises(ValueError):
        mv.CodeExtractor(vq, code_dtype=torch.float32)
    with pytest.raises(RuntimeError, match="CUDA"):
        mv.CodeExtractor(vq).push(torch.randn(1, 64, 2, 2))
    with pytest.raises(RuntimeError):
        mv.CodeExtractor(vq).usage_count()
