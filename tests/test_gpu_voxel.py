"""GPU parity: voxel pruning kernels (through the C-ABI) vs the CPU oracle, bit-exact."""
import numpy as np
import pytest
import torch

from oracle import voxel as ov
from streamvln_b200 import synth

pytestmark = pytest.mark.gpu


def _dev(st, cuda):
    d = {k: torch.from_numpy(st[k]).to(cuda) for k in ("depth", "pose", "intr")}
    d["feat"] = None if st["feat"] is None else torch.from_numpy(st["feat"].view(np.int16)).view(torch.bfloat16).to(cuda)
    return d


def _bits(t):
    return t.contiguous().view(torch.int16).cpu().numpy().view(np.uint16)


@pytest.mark.parametrize("S,T,G,D,vs,period,thr,pol", [
    (1, 32, 14, 3584, 0.2, 0, 0.0, 0),       # BASELINE config 1 shape
    (4, 32, 14, 256, 0.1, 0, 0.0, 1),
    (2, 9, 14, 64, 0.5, 4, 0.0, 1),          # ragged last period -> explicit segment offsets
    (1, 6, 27, 64, 0.1, 0, 0.4, 0),          # 729 tokens/frame + frame threshold
    (3, 5, 20, 64, 0.25, 2, 0.25, 1),
    (1, 1, 14, 64, 0.2, 0, 0.0, 0),
    (2, 48, 14, 64, 0.25, 0, 0.0, 0),        # 9408 tokens/segment > 8192 -> global hash table path
    (1, 16, 27, 64, 0.5, 0, 0.0, 1),         # 11664 tokens/segment, coarse voxels (many collisions)
    (2, 16, 24, 64, 0.2, 8, 0.0, 0),
])
def test_prune_matches_oracle(svln_lib, cuda, S, T, G, D, vs, period, thr, pol):
    from streamvln_b200.voxel import VoxelPruneConfig, VoxelPruner

    st = synth.make_streams(S, T, G, D, seed=S * 100 + T)
    cfg = VoxelPruneConfig(grid=G, dim=D, voxel_size=vs, period=period, frame_threshold=thr, invalid_policy=pol)
    want = ov.c_voxel_prune(st["depth"], st["pose"], st["intr"], st["feat"], G, vs, 10.0, period, cfg.min_keep, pol)
    pr = VoxelPruner(cfg, S, T, cuda)
    d = _dev(st, cuda)
    got = pr.prune(d["depth"], d["pose"], d["intr"], d["feat"])
    torch.cuda.synchronize()
    k = got.n_keep()
    assert np.array_equal(got.key.cpu().numpy(), want["key"])
    assert np.array_equal(got.mask.cpu().numpy(), want["mask"])
    assert k == want["n_keep"]
    assert np.array_equal(got.idx[:k].cpu().numpy(), want["idx"])
    assert np.array_equal(got.stream_keep_off.cpu().numpy(), want["stream_keep_off"])
    assert np.array_equal(_bits(got.feat[:k]), want["feat"])


def test_prune_edge_inputs(svln_lib, cuda):
    from streamvln_b200.voxel import VoxelPruneConfig, VoxelPruner

    G, T = 14, 4
    cfg = VoxelPruneConfig(grid=G, dim=32, voxel_size=0.2)
    pr = VoxelPruner(cfg, 1, T, cuda)
    st = synth.make_streams(1, T, G, 32, seed=5)
    # agent-style fake observation: depth = 0, pose = I (streamvln_agent.py:172-174): every token invalid
    st["depth"][:] = 0
    st["pose"][:] = np.eye(4, dtype=np.float32)
    for pol, expect in ((0, T * G * G), (1, 0)):
        cfg.invalid_policy = pol
        d = _dev(st, cuda)
        got = pr.prune(d["depth"], d["pose"], d["intr"], d["feat"])
        assert got.n_keep() == expect
    # nan / inf / negative / beyond max depth are all invalid; a static camera keeps only the last frame
    cfg.invalid_policy = 1
    st["depth"][:] = 2.0
    st["depth"][0, 0, :50] = np.nan
    st["depth"][0, 1, :50] = np.inf
    st["depth"][0, 2, :50] = -3.0
    st["depth"][0, 3, :50] = 12.0
    want = ov.c_voxel_prune(st["depth"], st["pose"], st["intr"], st["feat"], G, 0.2, 10.0, 0, 0, 1)
    d = _dev(st, cuda)
    got = pr.prune(d["depth"], d["pose"], d["intr"], d["feat"])
    assert np.array_equal(got.mask.cpu().numpy(), want["mask"])
    # identical poses + identical depth: frames 0..2 are fully superseded where frame 3 is valid
    m = want["mask"].reshape(T, G * G)
    assert m[:3].sum() <= 3 * 50 and m[3].sum() > 0


def test_keep_latest_ragged_segments(svln_lib, cuda):
    """Incremental slow-memory update shape: ragged segments, arbitrary (non-monotone) t."""
    from streamvln_b200.voxel import VoxelPruneConfig, VoxelPruner

    rng = np.random.default_rng(0)
    pr = VoxelPruner(VoxelPruneConfig(), 1, 1, cuda)
    for sizes in ([5], [1, 1, 1], [300, 0, 17, 4000], [9000, 50], [20000]):
        n = sum(sizes)
        seg = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
        key = rng.integers(0, max(2, n // 3), size=n).astype(np.int64)
        key[rng.random(n) < 0.05] = -1
        t = rng.integers(0, 40, size=n).astype(np.int32)
        for pol in (0, 1):
            want = ov.c_keep_latest(key, t, seg, pol)
            got = pr.keep_latest(torch.from_numpy(key).to(cuda), torch.from_numpy(t).to(cuda),
                                 torch.from_numpy(seg).to(cuda), max(sizes), pol)
            assert np.array_equal(got.cpu().numpy(), want), (sizes, pol)
            idx, n_keep, r = pr.compact(got, torch.from_numpy(seg).to(cuda))
            k = int(n_keep.item())
            assert np.array_equal(idx[:k].cpu().numpy(), np.nonzero(want)[0])
            assert np.array_equal(r.cpu().numpy(), np.concatenate([[0], np.cumsum([want[a:b].sum() for a, b in zip(seg[:-1], seg[1:])])]))


def test_full_size_properties(svln_lib, cuda):
    """BASELINE-size batch (64 streams x 32 frames x 196 tokens x 3584): size-independent properties."""
    from streamvln_b200.voxel import VoxelPruneConfig, VoxelPruner

    S, T, G, D = 64, 32, 14, 3584
    st = synth.make_streams(S, T, G, 8, seed=1, with_feat=False)
    cfg = VoxelPruneConfig(grid=G, dim=D, voxel_size=0.2, invalid_policy=1)
    pr = VoxelPruner(cfg, S, T, cuda)
    d = _dev(st, cuda)
    g = torch.Generator(device=cuda).manual_seed(0)
    feat = torch.randn(S, T, G * G, D, device=cuda, dtype=torch.bfloat16, generator=g)
    got = pr.prune(d["depth"], d["pose"], d["intr"], feat)
    k = got.n_keep()
    mask = got.mask.bool()
    idx = got.idx[:k].long()
    # sortedness, mask/idx agreement, per-stream offsets
    assert torch.all(idx[1:] > idx[:-1])
    assert torch.equal(torch.nonzero(mask).squeeze(1), idx)
    per_stream = mask.view(S, -1).sum(1)
    assert torch.equal(got.stream_keep_off[1:].long(), per_stream.cumsum(0))
    # every surviving key is unique within its stream and every valid key has exactly one survivor
    key = got.key.view(S, -1)
    for s in (0, 17, 63):
        ks = key[s][mask.view(S, -1)[s]]
        assert ks.unique().numel() == ks.numel()
        assert torch.equal(ks.unique(), key[s][key[s] >= 0].unique())
    # gather is a pure row copy: checksum of checksums
    assert torch.equal(got.feat[:k], feat.view(-1, D)[idx])
    # idempotence: pruning the survivors keeps all of them
    t_kept = ((idx // (G * G)) % T).int()
    seg = got.stream_keep_off.long()
    again = pr.keep_latest(key.reshape(-1)[idx].contiguous(), t_kept, seg, int(per_stream.max()), 1)
    assert bool(again.all())


def test_host_entry_matches_oracle(svln_lib, cuda):
    from streamvln_b200.voxel import VoxelPruneConfig, voxel_prune_host

    S, T, G, D = 2, 8, 14, 128
    st = synth.make_streams(S, T, G, D, seed=9)
    cfg = VoxelPruneConfig(grid=G, dim=D, voxel_size=0.2)
    want = ov.c_voxel_prune(st["depth"], st["pose"], st["intr"], st["feat"], G, 0.2)
    got = voxel_prune_host(cfg, st["depth"], st["pose"], st["intr"], st["feat"])
    assert got["n_keep"] == want["n_keep"]
    for k in ("mask", "idx", "stream_keep_off", "feat"):
        assert np.array_equal(got[k], want[k]), k


def test_splice_rows(svln_lib, cuda):
    from streamvln_b200 import _lib

    lib = svln_lib
    D = 256
    emb = torch.randn(100, D, device=cuda).bfloat16()
    feat = torch.randn(50, D, device=cuda).bfloat16()
    src = torch.tensor([3, 99, -1, -50, 0, -7], dtype=torch.int32, device=cuda)
    out = torch.empty(6, D, device=cuda, dtype=torch.bfloat16)
    _lib.check(lib.svln_splice_rows(emb.data_ptr(), feat.data_ptr(), src.data_ptr(), 6, D * 2, out.data_ptr(),
                                    _lib.cur_stream()))
    want = torch.stack([emb[3], emb[99], feat[0], feat[49], emb[0], feat[6]])
    assert torch.equal(out, want)


def test_error_codes(svln_lib, cuda):
    from streamvln_b200 import _lib

    lib = svln_lib
    assert lib.svln_gather_rows(None, 24, None, None, 1, None, None) == -22
    assert b"row_bytes" in lib.svln_last_error()
    x = torch.zeros(16, device=cuda)
    assert lib.svln_voxel_keys(x.data_ptr(), x.data_ptr(), x.data_ptr(), 1, 4, 4, 1, 0.0, 10.0, x.data_ptr(), None) == -22
    with pytest.raises(_lib.SvlnError):
        _lib.check(lib.svln_compact_mask(x.data_ptr(), 10, x.data_ptr(), 0, x.data_ptr(), x.data_ptr(), None, 0, None, None))
