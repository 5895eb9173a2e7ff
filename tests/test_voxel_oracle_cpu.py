"""CPU tests: the C oracle against the independent numpy restatement and hand-made cases."""
import numpy as np
import pytest

from oracle import voxel as ov
from streamvln_b200 import synth


def test_keys_c_vs_numpy_bit_exact():
    for seed, G, vs in [(0, 14, 0.2), (1, 27, 0.1), (2, 20, 0.5), (3, 24, 0.25)]:
        st = synth.make_stream(8, G, 8, seed, with_feat=False)
        a = ov.c_voxel_keys(st["depth"], st["pose"], st["intr"], G, vs)
        b = ov.np_voxel_keys(st["depth"], st["pose"], st["intr"], G, vs)
        assert a.dtype == np.int64 and a.shape == (8, G, G)
        assert np.array_equal(a, b)
        assert (a == -1).any() and (a >= 0).any()  # the 2 % depth holes produce invalid keys


def test_key_hand_case():
    # one frame, identity rotation, camera at origin: the centre token of a 1x1 grid looks along +z
    depth = np.full((1, 4, 4), 2.0, np.float32)
    pose = np.eye(4, dtype=np.float32)[None]
    K = np.array([[2, 0, 2, 0], [0, 2, 2, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)[None]
    key = ov.c_voxel_keys(depth, pose, K, 1, 0.5)
    # u = v = 2.0 -> x = y = 0, z = 2.0 -> voxel (0, 0, 4)
    off = 1 << 20
    assert key[0, 0, 0] == ((0 + off) << 42) | ((0 + off) << 21) | (4 + off)
    # invalid depth: zero, negative, nan, inf, >= max_depth
    for bad in (0.0, -1.0, np.nan, np.inf, 10.0, 11.0):
        depth[:] = bad
        assert ov.c_voxel_keys(depth, pose, K, 1, 0.5)[0, 0, 0] == -1
        assert ov.np_voxel_keys(depth, pose, K, 1, 0.5)[0, 0, 0] == -1


def test_negative_coordinates_floor():
    depth = np.full((1, 4, 4), 1.0, np.float32)
    pose = np.eye(4, dtype=np.float32)[None].copy()
    pose[0, 0, 3] = -0.05  # x_w = -0.05 -> floor(-0.05/0.2) = -1, not 0
    K = np.array([[2, 0, 2, 0], [0, 2, 2, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)[None]
    key = int(ov.c_voxel_keys(depth, pose, K, 1, 0.2)[0, 0, 0])
    assert (key >> 42) - (1 << 20) == -1


def test_keep_latest_semantics():
    #        idx: 0  1  2  3  4  5  6   7
    key = np.array([5, 7, 5, -1, 7, 5, 9, -1], np.int64)
    t = np.array([0, 0, 1, 1, 1, 1, 2, 2], np.int32)
    seg = np.array([0, 8], np.int64)
    keep = ov.c_keep_latest(key, t, seg, ov.INVALID_KEEP)
    # key 5: latest t=1 at idx 2 and 5 -> lowest index 2; key 7: idx 4; key 9: idx 6; invalid kept
    assert keep.tolist() == [0, 0, 1, 1, 1, 0, 1, 1]
    drop = ov.c_keep_latest(key, t, seg, ov.INVALID_DROP)
    assert drop.tolist() == [0, 0, 1, 0, 1, 0, 1, 0]
    # two segments: dedup does not cross the boundary
    seg2 = np.array([0, 4, 8], np.int64)
    assert ov.c_keep_latest(key, t, seg2, ov.INVALID_DROP).tolist() == [0, 1, 1, 0, 1, 1, 1, 0]
    for pol in (0, 1):
        for s in (seg, seg2):
            assert np.array_equal(ov.c_keep_latest(key, t, s, pol), ov.np_keep_latest(key, t, s, pol))


@pytest.mark.parametrize("S,T,G,vs,period,thr,pol", [
    (1, 32, 14, 0.2, 0, 0, 0), (2, 9, 14, 0.5, 4, 0, 1), (1, 6, 27, 0.1, 0, 300, 0), (3, 5, 20, 0.25, 2, 100, 1),
    (1, 1, 14, 0.2, 0, 0, 0), (2, 3, 24, 0.25, 5, 0, 0),
])
def test_prune_c_vs_numpy(S, T, G, vs, period, thr, pol):
    st = synth.make_streams(S, T, G, 16, seed=7)
    a = ov.c_voxel_prune(st["depth"], st["pose"], st["intr"], st["feat"], G, vs, 10.0, period, thr, pol)
    b = ov.np_voxel_prune(st["depth"], st["pose"], st["intr"], st["feat"], G, vs, 10.0, period, thr, pol)
    for k in ("key", "mask", "idx", "stream_keep_off", "feat"):
        assert np.array_equal(a[k], b[k]), k
    assert a["n_keep"] == b["n_keep"] == int(a["mask"].sum())
    assert np.all(np.diff(a["idx"]) > 0)


def test_prune_empty():
    z = np.zeros((0, 0, 384, 384), np.float32)
    a = ov.c_voxel_prune(z, np.zeros((0, 0, 4, 4)), np.zeros((0, 0, 4, 4)), None, 14, 0.2)
    assert a["n_keep"] == 0 and a["idx"].size == 0


def test_torch_reference_idiom_matches_oracle():
    import torch

    st = synth.make_stream(12, 14, 32, seed=3)
    want = ov.c_voxel_prune(st["depth"][None], st["pose"][None], st["intr"][None], st["feat"][None], 14, 0.2)
    feat = torch.from_numpy(st["feat"].view(np.int16)).view(torch.bfloat16)
    mask, kept = ov.torch_reference_idiom_prune(torch.from_numpy(st["depth"]), torch.from_numpy(st["pose"]),
                                                torch.from_numpy(st["intr"]), feat, 14, 0.2)
    assert np.array_equal(mask.numpy().astype(np.uint8), want["mask"])
    assert np.array_equal(kept.view(torch.int16).numpy().view(np.uint16), want["feat"])


def test_idempotent_and_monotone():
    # pruning the survivors again keeps all of them; a coarser voxel never keeps more tokens
    st = synth.make_stream(16, 14, 8, seed=11)
    k1 = ov.c_voxel_prune(st["depth"][None], st["pose"][None], st["intr"][None], None, 14, 0.2)
    key = k1["key"].reshape(-1)[k1["idx"]]
    t = (k1["idx"] // 196).astype(np.int32)
    again = ov.c_keep_latest(key, t, np.array([0, key.size], np.int64))
    assert again.all()
    k2 = ov.c_voxel_prune(st["depth"][None], st["pose"][None], st["intr"][None], None, 14, 0.4)
    assert k2["n_keep"] <= k1["n_keep"]
