"""The eager TF-1 stand-in (`oracle/tf1_eager.py`, test infrastructure) checked against TensorFlow's DOCUMENTED
behaviour: the worked examples of the TF-1.x API docs (tf.pad, tf.scatter_nd, tf.gather_nd, variable_scope naming)
and brute-force loops written from the op definitions (conv2d SAME/VALID, conv2d_transpose as the scatter form of
the conv2d input-gradient, pooling).  The golden vectors of tests/golden/ref_*.npz are only as good as this layer."""
import numpy as np
import pytest
import torch

from oracle import tf1_eager as E


def T(a):
    return E.Tensor(torch.as_tensor(np.asarray(a)))


def test_pad_examples_of_the_tf_docs():
    t = T(np.array([[1, 2, 3], [4, 5, 6]], dtype=np.float64))
    const = E.pad(t, [[1, 1], [2, 2]], "CONSTANT").numpy()
    assert const.tolist() == [[0, 0, 0, 0, 0, 0, 0], [0, 0, 1, 2, 3, 0, 0], [0, 0, 4, 5, 6, 0, 0], [0, 0, 0, 0, 0, 0, 0]]
    refl = E.pad(t, [[1, 1], [2, 2]], "REFLECT").numpy()
    assert refl.tolist() == [[6, 5, 4, 5, 6, 5, 4], [3, 2, 1, 2, 3, 2, 1], [6, 5, 4, 5, 6, 5, 4], [3, 2, 1, 2, 3, 2, 1]]


def test_scatter_nd_and_gather_nd_examples_of_the_tf_docs():
    out = E.scatter_nd(T([[4], [3], [1], [7]]), T(np.array([9., 10., 11., 12.])), [8]).numpy()
    assert out.tolist() == [0, 11, 0, 10, 9, 0, 0, 12]
    dup = E.scatter_nd(T([[1], [1]]), T(np.array([2., 3.])), [3]).numpy()          # duplicates accumulate
    assert dup.tolist() == [0, 5, 0]
    params = T(np.arange(24, dtype=np.float64).reshape(2, 3, 4))
    got = E.gather_nd(params, T(np.array([[[2], [0]], [[1], [1]]])), batch_dims=1).numpy()
    assert got.tolist() == [params.numpy()[0, [2, 0]].tolist(), params.numpy()[1, [1, 1]].tolist()]


def test_same_padding_rule():
    # TF: out = ceil(n / s); total = max((out - 1) s + k - n, 0); before = total // 2 (the extra sample goes after)
    assert E._same_pads(8, 3, 1) == (1, 1)
    assert E._same_pads(8, 3, 2) == (0, 1)
    assert E._same_pads(9, 3, 2) == (1, 1)
    assert E._same_pads(8, 4, 2) == (1, 1)
    assert E._same_pads(5, 1, 1) == (0, 0)
    assert E._same_pads(7, 6, 2) == (2, 3)


def _conv2d_loops(x, w, s, padding):
    B, H, W, Ci = x.shape
    kh, kw, _, Co = w.shape
    if padding == "SAME":
        pt, pb = E._same_pads(H, kh, s)
        pl, pr = E._same_pads(W, kw, s)
    else:
        pt = pb = pl = pr = 0
    xp = np.zeros((B, H + pt + pb, W + pl + pr, Ci))
    xp[:, pt:pt + H, pl:pl + W] = x
    Ho, Wo = (xp.shape[1] - kh) // s + 1, (xp.shape[2] - kw) // s + 1
    y = np.zeros((B, Ho, Wo, Co))
    for i in range(Ho):
        for j in range(Wo):
            patch = xp[:, i * s:i * s + kh, j * s:j * s + kw, :]
            y[:, i, j, :] = np.einsum("bhwc,hwco->bo", patch, w)
    return y, (pt, pl)


@pytest.mark.parametrize("H,k,s,padding", [(8, 3, 1, "SAME"), (8, 3, 2, "SAME"), (9, 3, 2, "SAME"), (10, 3, 2, "VALID"),
                                           (8, 4, 2, "SAME"), (6, 1, 1, "VALID")])
def test_conv2d_against_loops(H, k, s, padding):
    rng = np.random.default_rng(0)
    x, w = rng.standard_normal((2, H, H, 3)), rng.standard_normal((k, k, 3, 5))
    want, _ = _conv2d_loops(x, w, s, padding)
    got = E.conv2d(T(x), T(w), [1, s, s, 1], padding).numpy()
    assert got.shape == want.shape and np.abs(got - want).max() < 1e-12


@pytest.mark.parametrize("H,k,s", [(4, 4, 2), (4, 3, 1), (4, 3, 2), (3, 6, 2), (5, 4, 2), (4, 4, 1)])
def test_conv2d_transpose_is_the_scatter_form_of_the_conv2d_input_gradient(H, k, s):
    """conv2d_transpose(value, filter[kh,kw,out,in], output_shape, SAME): every input sample (i, j) adds
    value[i,j,:] . filter[ky,kx,o,:] at output (i*s + ky - pad_top, j*s + kx - pad_left) when inside."""
    rng = np.random.default_rng(1)
    Ci, Co = 3, 4
    x, w = rng.standard_normal((2, H, H, Ci)), rng.standard_normal((k, k, Co, Ci))
    Ho = H * s
    pt, _ = E._same_pads(Ho, k, s)
    want = np.zeros((2, Ho, Ho, Co))
    for i in range(H):
        for j in range(H):
            for ky in range(k):
                for kx in range(k):
                    oy, ox = i * s + ky - pt, j * s + kx - pt
                    if 0 <= oy < Ho and 0 <= ox < Ho:
                        want[:, oy, ox, :] += x[:, i, j, :] @ w[ky, kx].T
    got = E.conv2d_transpose(T(x), T(w), [2, Ho, Ho, Co], [1, s, s, 1], "SAME").numpy()
    assert np.abs(got - want).max() < 1e-12


def test_conv2d_transpose_is_differentiable_in_value_and_filter():
    x = torch.randn(1, 3, 3, 2, dtype=torch.float64, requires_grad=True)
    w = torch.randn(4, 4, 3, 2, dtype=torch.float64, requires_grad=True)
    f = lambda a, b: E.conv2d_transpose(E.Tensor(a), E.Tensor(b), [1, 6, 6, 3], [1, 2, 2, 1], "SAME").t
    assert torch.autograd.gradcheck(f, (x, w), eps=1e-6, atol=1e-6)


def test_pooling_moments_batchnorm_l2norm_xent_formulas():
    rng = np.random.default_rng(2)
    x = rng.standard_normal((2, 4, 6, 3))
    mp = E.max_pooling2d(T(x), 2, 2, "SAME").numpy()
    ap = E.average_pooling2d(T(x), 2, 2, "SAME").numpy()
    for i in range(2):
        for j in range(3):
            win = x[:, 2 * i:2 * i + 2, 2 * j:2 * j + 2, :]
            assert np.allclose(mp[:, i, j], win.max(axis=(1, 2))) and np.allclose(ap[:, i, j], win.mean(axis=(1, 2)))
    m, v = E.moments(T(x), [0, 1, 2])
    assert np.allclose(m.numpy(), x.mean(axis=(0, 1, 2))) and np.allclose(v.numpy(), x.var(axis=(0, 1, 2)))       # biased
    y = E.batch_normalization(T(x), m, v, T(np.full(3, 0.5)), T(np.full(3, 2.0)), 1e-5).numpy()
    assert np.allclose(y, (x - x.mean(axis=(0, 1, 2))) / np.sqrt(x.var(axis=(0, 1, 2)) + 1e-5) * 2.0 + 0.5)
    a = rng.standard_normal((1, 7))
    assert np.allclose(E.l2_normalize(T(a)).numpy(), a / np.sqrt((a * a).sum()))
    assert np.allclose(E.l2_normalize(T(np.zeros((1, 3)))).numpy(), 0.0)                                            # eps guard
    z, lab = rng.standard_normal(9) * 4, (rng.random(9) > 0.5).astype(np.float64)
    want = -(lab * np.log(1 / (1 + np.exp(-z))) + (1 - lab) * np.log(1 - 1 / (1 + np.exp(-z))))
    assert np.allclose(E.sigmoid_cross_entropy_with_logits(labels=T(lab), logits=T(z)).numpy(), want)
    assert float(E.l2_loss(T(np.array([3.0, 4.0]))).numpy()) == 12.5


def test_variable_scope_naming_reuse_and_deferred_assign():
    E.STATE.reset()
    E.STATE.init_fn = lambda full, shape, kind, value, stddev: np.full(shape, 1.0 if kind != "const" else value, np.float32)
    E.STATE.initial_values = None
    with E.variable_scope("discriminator"):
        with E.variable_scope("D_logit") as sc:
            assert sc.name == "discriminator/D_logit"
            a = E.layers_dense(T(np.ones((2, 3))), 4, kernel_initializer=E.truncated_normal_initializer(stddev=0.02))
        with E.variable_scope(None, default_name="prelu"):
            p0 = E.get_variable("alpha", [2], initializer=E.constant_initializer(0.0))
        with E.variable_scope(None, default_name="prelu"):
            p1 = E.get_variable("alpha", [2], initializer=E.constant_initializer(0.0))
    assert p0.name == "discriminator/prelu/alpha:0" and p1.name == "discriminator/prelu_1/alpha:0"
    assert "discriminator/D_logit/dense/kernel" in E.STATE.vars and a.shape.as_list() == [2, 4]
    with pytest.raises(ValueError):                       # second creation without reuse
        with E.variable_scope("discriminator"):
            with E.variable_scope("D_logit"):
                E.get_variable("dense/kernel", [3, 4])
    with E.variable_scope("discriminator", reuse=True):   # reuse: the layer is called 'dense' again, same variable
        with E.variable_scope("D_logit"):
            E.layers_dense(T(np.ones((2, 3))), 4, kernel_initializer=E.truncated_normal_initializer(stddev=0.02))
            with pytest.raises(ValueError):
                E.get_variable("missing", [1])
    assert sum(k.startswith("discriminator/D_logit/") for k in E.STATE.vars) == 2
    v = E.STATE.vars["discriminator/prelu/alpha"]
    v.assign(T(np.array([5.0, 6.0])))
    assert v.numpy().tolist() == [0.0, 0.0]               # reads inside the run still see the old value
    E.end_run()
    assert v.numpy().tolist() == [5.0, 6.0]


def test_more_examples_of_the_tf_docs():
    # tf.nn.depth_to_space: [1,1,1,4] with block 2 -> [[[[1],[2]],[[3],[4]]]]; the [1,1,1,12] example -> 2x2 RGB pixels
    x = T(np.array([[[[1., 2., 3., 4.]]]]))
    assert E.depth_to_space(x, 2).numpy().tolist() == [[[[1.0], [2.0]], [[3.0], [4.0]]]]
    x = T(np.arange(1., 13.).reshape(1, 1, 1, 12))
    assert E.depth_to_space(x, 2).numpy().tolist() == [[[[1., 2., 3.], [4., 5., 6.]], [[7., 8., 9.], [10., 11., 12.]]]]
    # tf.split with sizes, tf.meshgrid(indexing='ij'), tf.clip_by_value, tf.cast float -> int truncation, floor div / mod on ints
    parts = E.split(T(np.arange(10.).reshape(1, 10)), [4, 6], axis=-1)
    assert [p.shape.as_list() for p in parts] == [[1, 4], [1, 6]]
    a, b = E.meshgrid(E.range_(2), E.range_(3), indexing="ij")
    assert a.numpy().tolist() == [[0, 0, 0], [1, 1, 1]] and b.numpy().tolist() == [[0, 1, 2], [0, 1, 2]]
    assert E.clip_by_value(T(np.array([-3, 0, 9])), 0, 5).numpy().tolist() == [0, 0, 5]
    assert E.cast(T(np.array([8.5, -1.5, 16.99])), E.int32).numpy().tolist() == [8, -1, 16]
    seven = T(np.array(7))
    assert int(seven // 2) == 3 and int(seven % 2) == 1 and int(1 - seven % 2) == 0
    # tf.nn.leaky_relu, tf.image.resize_nearest_neighbor, tf.reduce_mean(keepdims), tf.norm
    assert E.leaky_relu(T(np.array([-2., 3.])), 0.2).numpy().tolist() == [-0.4, 3.0]
    up = E.resize_nearest_neighbor(T(np.arange(4.).reshape(1, 2, 2, 1)), [4, 4]).numpy()[0, :, :, 0]
    assert up.tolist() == [[0, 0, 1, 1], [0, 0, 1, 1], [2, 2, 3, 3], [2, 2, 3, 3]]
    m = E.reduce_mean(T(np.arange(24.).reshape(2, 3, 4)), axis=[1, 2], keepdims=True)
    assert m.shape.as_list() == [2, 1, 1] and m.numpy().ravel().tolist() == [5.5, 17.5]
    assert float(E.norm(T(np.array([3., 4.]))).numpy()) == 5.0
