"""The torch-CPU oracle of the callers either side of the rasterizer (oracle/slam_oracle.py) against
independent formulations -- runs without a GPU."""
import numpy as np
import torch

from oracle import slam_oracle as SO


def test_pose_matrix_matches_scipy_and_transform():
    from scipy.spatial.transform import Rotation
    g = torch.Generator().manual_seed(1)
    q, t = torch.randn(4, generator=g, dtype=torch.float64), torch.randn(3, generator=g, dtype=torch.float64)
    m = SO.pose_matrix(q, t)
    qn = (q / q.norm()).numpy()
    want = Rotation.from_quat([qn[1], qn[2], qn[3], qn[0]]).as_matrix()         # scipy: (x, y, z, w)
    assert np.allclose(m[:3, :3].numpy(), want, atol=1e-12) and np.allclose(m[:3, 3].numpy(), t.numpy())
    assert np.allclose(m[3].numpy(), [0, 0, 0, 1])
    pts = torch.randn(17, 3, generator=g, dtype=torch.float64)
    assert torch.allclose(SO.transform_points(m, pts), pts @ m[:3, :3].T + m[:3, 3], atol=1e-12)


def test_ssim_matches_a_direct_numpy_convolution():
    g = torch.Generator().manual_seed(2)
    a, b = torch.rand(2, 13, 17, generator=g, dtype=torch.float64), torch.rand(2, 13, 17, generator=g, dtype=torch.float64)
    w1 = SO.gaussian_window().double().numpy()
    assert abs(w1.sum() - 1) < 1e-6 and w1.argmax() == 5 and np.allclose(w1, w1[::-1])

    def conv(img):                                      # zero-padded 11x11 correlation, pixel by pixel
        C, H, W = img.shape
        pad = np.zeros((C, H + 10, W + 10))
        pad[:, 5:-5, 5:-5] = img
        out = np.zeros_like(img)
        for y in range(H):
            for x in range(W):
                out[:, y, x] = (pad[:, y:y + 11, x:x + 11] * np.outer(w1, w1)).sum((1, 2))
        return out
    x, y = a.numpy(), b.numpy()
    mu1, mu2 = conv(x), conv(y)
    s1, s2, s12 = conv(x * x) - mu1 ** 2, conv(y * y) - mu2 ** 2, conv(x * y) - mu1 * mu2
    want = (((2 * mu1 * mu2 + 1e-4) * (2 * s12 + 9e-4)) / ((mu1 ** 2 + mu2 ** 2 + 1e-4) * (s1 + s2 + 9e-4))).mean()
    assert abs(float(SO.ssim(a, b)) - want) < 1e-7          # the window products are rounded to float32 (utils/slam_external.py:62)
    assert abs(float(SO.ssim(a, a)) - 1.0) < 1e-12
    assert abs(float(SO.mapping_colour_loss(a, b)) - (0.8 * np.abs(x - y).mean() + 0.2 * (1 - want))) < 1e-7


def test_tree_losses_match_a_hand_written_log_softmax():
    g = torch.Generator().manual_seed(3)
    sizes, L, H, W = [3, 4, 2], 7, 5, 6
    S = sum(sizes)
    sem = torch.randn(S, H, W, generator=g, dtype=torch.float64)
    labels = torch.stack([torch.randint(0, n, (H, W), generator=g) for n in sizes + [L]])
    weight = torch.randn(L, S, generator=g, dtype=torch.float64)
    bias = torch.randn(L, generator=g, dtype=torch.float64)
    want, beg = 0.0, 0
    for l, n in enumerate(sizes):
        z = sem[beg:beg + n].reshape(n, -1)
        lse = torch.logsumexp(z, 0)
        want += float((lse - z.gather(0, labels[l].reshape(1, -1))[0]).mean())
        beg += n
    assert abs(float(SO.level_cross_entropy(sem, labels, sizes)) - want) < 1e-12
    z = weight @ sem.reshape(S, -1) + bias[:, None]
    leaf = float((torch.logsumexp(z, 0) - z.gather(0, labels[-1].reshape(1, -1))[0]).mean())
    assert abs(float(SO.leaf_cross_entropy(sem, labels[-1], weight, bias)) - leaf) < 1e-12
    assert abs(float(SO.tree_semantic_loss(sem, labels, sizes, weight, bias)) - (want + 5 * leaf)) < 1e-11
    lab2 = labels.clone()
    lab2[0, 0] = -100                                  # ignored pixels drop out of the mean
    z0 = sem[:3].reshape(3, -1)[:, W:]
    keep = float((torch.logsumexp(z0, 0) - z0.gather(0, labels[0].reshape(1, -1)[:, W:])[0]).mean())
    assert abs(float(SO.level_cross_entropy(sem, lab2, sizes[:1])) - keep) < 1e-12


def test_tracking_loss_mask_and_weights():
    g = torch.Generator().manual_seed(4)
    H, W = 6, 7
    im, gt_im = torch.rand(3, H, W, generator=g), torch.rand(3, H, W, generator=g)
    depth, gt_depth = 1 + torch.rand(1, H, W, generator=g), 1 + torch.rand(1, H, W, generator=g)
    sil = torch.rand(1, H, W, generator=g)
    gt_depth[0, 0, :3] = 0
    depth[0, 1, 0] = float("nan")
    want = 0.0
    for y in range(H):
        for x in range(W):
            if gt_depth[0, y, x] > 0 and not torch.isnan(depth[0, y, x]) and sil[0, y, x] > 0.5:
                want += float(abs(gt_depth[0, y, x] - depth[0, y, x])) + 0.5 * float((gt_im[:, y, x] - im[:, y, x]).abs().sum())
    assert abs(float(SO.tracking_loss(im, depth, sil, gt_im, gt_depth, sil_thres=0.5)) - want) < 1e-4
    assert float(SO.tracking_loss(im, depth, sil, gt_im, gt_depth, sil_thres=0.5, use_sil_for_loss=False)) > want


def test_adam_step_matches_torch_optim_adam():
    g = torch.Generator().manual_seed(5)
    p0 = torch.randn(50, 3, generator=g)
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=2.5e-3, eps=1e-15)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for step in range(1, 5):
        grad = torch.randn(50, 3, generator=g)
        ref.grad = grad.clone()
        opt.step()
        SO.adam_step(p, grad, m, v, step, 2.5e-3, eps=1e-15)
        assert torch.allclose(p, ref.detach(), rtol=1e-6, atol=1e-9)
    st = opt.state[ref]
    assert torch.allclose(m, st["exp_avg"], rtol=1e-6, atol=1e-12) and torch.allclose(v, st["exp_avg_sq"], rtol=1e-6, atol=1e-12)


def test_prune_mask_and_remove_points():
    lo = torch.tensor([[-6.0], [0.0], [3.0], [-5.2]])
    ls = torch.log(torch.tensor([[0.01], [0.5], [0.02], [0.01]]))
    assert SO.prune_mask(lo, ls, 0.005, 1.0, False).tolist() == [True, False, False, False]
    assert SO.prune_mask(lo, ls, 0.006, 1.0, True).tolist() == [True, True, False, True]
    out = SO.remove_points({"a": torch.arange(8.0).view(4, 2), "b": torch.arange(4)}, torch.tensor([True, False, False, True]))
    assert out["a"].tolist() == [[2.0, 3.0], [4.0, 5.0]] and out["b"].tolist() == [1, 2]


def test_keyframe_overlap_counts_match_numpy_loops():
    g = torch.Generator().manual_seed(6)
    W, H = 64, 48
    K = torch.tensor([[50.0, 0, 31.5], [0, 50.0, 23.5], [0, 0, 1]])
    z = 1 + 3 * torch.rand(200, generator=g)
    pts = torch.stack(((torch.rand(200, generator=g) * W - 31.5) * z / 50, (torch.rand(200, generator=g) * H - 23.5) * z / 50, z), 1)
    w2cs = torch.eye(4).repeat(3, 1, 1)
    w2cs[1, 0, 3] = 0.4
    w2cs[2] = torch.diag(torch.tensor([-1.0, 1, -1, 1]))
    want = []
    for m in w2cs.numpy():
        c = 0
        for p in pts.numpy():
            X = m[:3, :3] @ p + m[:3, 3]
            pz = X[2] + 1e-5
            u, v = (50 * X[0] + 31.5 * X[2]) / pz, (50 * X[1] + 23.5 * X[2]) / pz
            c += int(20 < u < W - 20 and 20 < v < H - 20 and pz > 0)
        want.append(c)
    got = SO.keyframe_overlap_counts(pts, w2cs, K, W, H)
    assert all(abs(a - b) <= 1 for a, b in zip(got, want)) and got[2] == 0 and got[0] > 10


def test_backproject_samples_reproduces_the_reference_duplicate_removal():
    """hier_slam_b200.keyframes.backproject_samples (host torch code, runs on CPU) == get_pointcloud of
    utils/keyframe_selection.py:10-37 restated step by step, including its removal of every duplicated row."""
    from hier_slam_b200.keyframes import backproject_samples
    g = torch.Generator().manual_seed(7)
    H, W = 12, 16
    K = torch.tensor([[20.0, 0, 7.5], [0, 20.0, 5.5], [0, 0, 1]])
    depth = 1 + torch.rand(1, H, W, generator=g)
    depth[0, 3, 4] = 0.0                                   # a zero-depth pixel back-projects onto the camera centre
    w2c = torch.eye(4)                                     # camera at the world origin: that point IS (0, 0, 0)
    idx = torch.tensor([[3, 4], [2, 2], [5, 9], [2, 2], [7, 1], [11, 15]])       # (2, 2) drawn twice
    got = backproject_samples(depth, K, w2c, idx)
    z = depth[0, idx[:, 0], idx[:, 1]]
    want = torch.stack(((idx[:, 1] - 7.5) / 20 * z, (idx[:, 0] - 5.5) / 20 * z, z), 1)[[2, 4, 5]]
    assert got.shape == (3, 3) and torch.allclose(got, want, atol=1e-6)
