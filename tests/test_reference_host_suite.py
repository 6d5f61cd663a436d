"""The host-side behaviours the reference's own test-suite pins for the parameter tree and the driver, restated against
the mirrored API: testing/test_param.py (naming, tree surgery, assignment, LOCAL feed layout, ParamList, initialisation
range), testing/test_data.py (Data values / replacement), testing/test_model.py (keyword set-up, Indexer set-up at
compile time is a GPU test: tests/test_gpu_api.py).  No GPU and no kernel: everything here is the Python tree that
sits above the C ABI.  Each test names the upstream test it restates."""
import numpy as np
import pytest
import torch

import henbun_b200 as hb
from henbun_b200.param import graph_key


# ---------------------------------------------------------------------------------------------- test_param.py:11-28
def test_unnamed_variable():
    assert hb.param.Variable([1]).name == 'unnamed'


def test_parent_that_does_not_hold_the_child_is_an_error():
    p = hb.param.Variable([1])
    p._parent = hb.param.Parameterized()                 # "do not do this" upstream
    with pytest.raises(ValueError):
        p.name


def test_two_names_for_one_child_is_an_error():
    m = hb.param.Parameterized()
    m.p = hb.param.Variable([1])
    m.p2 = m.p
    with pytest.raises(ValueError):
        m.p.name


# ---------------------------------------------------------------------------------------------- test_param.py:30-124
@pytest.fixture
def scalar_model():
    m = hb.model.Model()
    m.p = hb.param.Variable([1], transform=hb.transforms.positive)
    m.q = hb.param.Variable([3], collections=graph_key.LOCAL)
    m.r = hb.param.Variable([5, 4], collections=graph_key.LOCAL)
    m.s = hb.param.Variable([5, 4], n_batch=2)
    return m


def test_sorted_variables(scalar_model):
    m = scalar_model
    assert m.sorted_variables[0] is m.p and len(m.sorted_variables) == 4


def test_collections_select_variables(scalar_model):
    m = scalar_model
    glob = m.get_variables(graph_key.VARIABLES)
    assert any(v is m.p for v in glob) and not any(v is m.q for v in glob)
    loc = m.get_variables(graph_key.LOCAL)
    assert any(v is m.q for v in loc) and any(v is m.r for v in loc)


def test_assignment_keeps_the_variable_and_applies_at_value(scalar_model):
    m = scalar_model
    assert m.p._assigned                                  # a fresh Variable carries its initial value as a pending assignment
    before = m.p.value
    assert not m.p._assigned                              # .value applied it (test_param.py:64-71)
    m.p = 3.0
    assert m.p._assigned and isinstance(m.p, hb.param.Variable)
    after = m.p.value
    assert not m.p._assigned
    assert not np.allclose(before, [3.0]) and np.allclose(after, [3.0], atol=1e-5)


def test_replacing_a_child_rewires_the_parents(scalar_model):
    m = scalar_model
    old, new = m.p, hb.param.Variable(3)
    m.p = new
    assert new.highest_parent is m and old.highest_parent is not m
    assert m.p is new and new.name == 'p'


def test_feed_size_counts_the_local_variables(scalar_model):
    assert scalar_model.feed_size == 3 + 5 * 4


def test_variable_feed(scalar_model):
    m = scalar_model
    vq, vr = np.ones((10, 3), np.float32), np.ones((10, 20), np.float32)
    m.q.feed(torch.from_numpy(vq)); m.r.feed(torch.from_numpy(vr))
    m.s = np.ones((2, 5, 4))
    assert np.allclose(m.q.value, vq) and np.allclose(m.r.value.ravel(), vr.ravel()) and np.allclose(m.s.value, 1.0)
    assert m.r.value.shape == (10, 5, 4)


def test_parameterized_feed_splits_in_name_order(scalar_model):
    m = scalar_model
    val = np.random.RandomState(0).randn(10, m.feed_size).astype(np.float32)
    m.feed(torch.from_numpy(val))
    assert np.allclose(m.q.value, val[:, :3]) and np.allclose(m.r.value.ravel(), val[:, 3:].ravel())


# ---------------------------------------------------------------------------------------------- test_param.py:126-149
def test_layered_variables():
    m = hb.model.Model()
    m.p = hb.param.Variable([3], n_layers=[2, 3])
    m.q = hb.param.Variable([3], n_layers=[2, 3], collections=graph_key.LOCAL)
    m.r = hb.param.Variable([5, 4], n_layers=[2, 3], collections=graph_key.LOCAL)
    rng = np.random.RandomState(0)
    val = rng.randn(2, 3, 3)
    m.p = val
    assert np.allclose(m.p.value, val, atol=1e-6)
    fed = rng.randn(2, 3, 10, m.feed_size).astype(np.float32)
    m.feed(torch.from_numpy(fed))
    assert np.allclose(m.q.value.ravel(), fed[..., :3].ravel()) and np.allclose(m.r.value.ravel(), fed[..., 3:].ravel())
    assert m.r.value.shape == (2, 3, 10, 5, 4)


# ---------------------------------------------------------------------------------------------- test_param.py:152-202
@pytest.fixture
def deep_model():
    m = hb.model.Model()
    m.foo = hb.param.Parameterized()
    m.foo.bar = hb.param.Parameterized()
    m.foo.bar.baz = hb.param.Variable(1)
    m.foo.bar.q = hb.param.Variable([3], collections=graph_key.LOCAL)
    m.foo.bar.r = hb.param.Variable([5, 4], collections=graph_key.LOCAL)
    return m


def test_deep_tree_parents_and_names(deep_model):
    m = deep_model
    assert m.foo.highest_parent is m and m.foo.bar.highest_parent is m and m.foo.bar.baz.highest_parent is m
    assert (m.foo.name, m.foo.bar.name, m.foo.bar.baz.name) == ('foo', 'bar', 'baz')
    assert m.foo.bar.baz.long_name == 'model.foo.bar.baz'


def test_deep_tree_replacement(deep_model):
    m = deep_model
    old = m.foo.bar.baz
    m.foo.bar.baz = hb.param.Variable(3)
    assert old.highest_parent is not m
    old_foo, new = m.foo, hb.param.Variable(3)
    m.foo = new
    assert new.highest_parent is m and old_foo.highest_parent is not m


def test_deep_tree_assignment_and_feed(deep_model):
    m = deep_model
    before = m.foo.bar.baz.value
    m.foo.bar.baz = 3.0
    assert isinstance(m.foo.bar.baz, hb.param.Variable)
    assert not np.allclose(before, [3.0]) and np.allclose(m.foo.bar.baz.value, [3.0])
    val = np.random.RandomState(0).randn(10, m.feed_size).astype(np.float32)
    m.feed(torch.from_numpy(val))
    assert np.allclose(m.foo.bar.q.value.ravel(), val[:, :3].ravel())
    assert np.allclose(m.foo.bar.r.value.ravel(), val[:, 3:].ravel())


# ---------------------------------------------------------------------------------------------- test_param.py:205-266
def test_paramlist_construction_naming_and_membership():
    hb.param.ParamList([])
    with pytest.raises(AssertionError):
        hb.param.ParamList([hb.param.Variable(1), 'stringsnotallowed'])
    p1, p2 = hb.param.Variable([2]), hb.param.Variable([3, 5])
    lst = hb.param.ParamList([p1, p2])
    assert (p1.name, p2.name) == ('item0', 'item1')
    assert any(v is p1 for v in lst.sorted_variables) and any(v is p2 for v in lst.sorted_variables)
    p3 = hb.param.Variable([2, 3])
    lst.append(p3)
    assert any(v is p3 for v in lst.sorted_variables)
    with pytest.raises(AssertionError):
        lst.append('foo')


def test_paramlist_setitem_assigns_values_only():
    p1, p2 = hb.param.Variable([1]), hb.param.Variable([2, 3])
    m = hb.model.Model()
    m.l = hb.param.ParamList([p1, p2])
    m.l[0] = 1.0
    assert np.allclose(p1.value, 1.0)
    with pytest.raises(TypeError):
        m.l[0] = hb.param.Variable(12)


def test_paramlist_of_parameterized_follows_tf_mode():
    pzd = hb.param.Parameterized()
    p = hb.param.Variable([1])
    pzd.p = p
    m = hb.model.Model()
    m.l = hb.param.ParamList([pzd])
    m.l[0].p = 5
    assert np.allclose(p.value, 5.0)
    assert not pzd._tf_mode
    with m.tf_mode():
        assert pzd._tf_mode
    assert not pzd._tf_mode


# ---------------------------------------------------------------------------------------------- test_param.py:286-296
def test_initial_values_lie_within_two_standard_deviations():
    m = hb.model.Model()
    m.p = hb.param.Variable(shape=[10, 20], mean=0.0, stddev=0.1)
    m.q = hb.param.Variable(shape=[10, 20], mean=0.1, stddev=0.1)
    assert np.all(np.abs(m.p.value) < 0.2) and np.all(m.q.value > -0.1) and np.all(m.q.value < 0.3)


# ---------------------------------------------------------------------------------------------- test_param.py:298-316
def test_feed_dict_of_the_tree():
    """Upstream feeds data[index] from the host; here a MinibatchData feed carries the INDEX (the rows are gathered on the
    device from the resident array) and a Data feed carries the array."""
    rng = np.random.RandomState(0)
    m = hb.model.Model()
    m.foo = hb.param.Parameterized()
    m.foo.bar = hb.param.Parameterized()
    m.foo.bar.baz = hb.param.Variable(1)
    data, mb = rng.randn(3, 2), rng.randn(10, 3, 2)
    m.foo.bar.d1 = hb.param.Data(data)
    m.foo.bar.d2 = hb.param.MinibatchData(mb)
    index = rng.randint(0, 10, 3)
    fd = m.get_feed_dict(index)
    assert fd[m.foo.bar.d1] is data and np.array_equal(fd[m.foo.bar.d2], index)
    assert np.allclose(m.foo.bar.d2.data[fd[m.foo.bar.d2]], mb[index])
    assert m.foo.bar.d2.data_size == 10 and list(m.foo.bar.d2.shape) == [3, 2]


# ---------------------------------------------------------------------------------------------- test_data.py
def test_data_values_and_replacement():
    rng = np.random.RandomState(0)
    x, y, z = rng.randn(3, 2), rng.randn(4, 3), rng.randint(1, 30, 20)
    m = hb.model.Model()
    m.p = hb.param.Parameterized()
    m.x = hb.param.Data(x); m.p.y = hb.param.Data(y); m.p.z = hb.param.Data(z)
    assert np.allclose(m.x.value, x) and np.allclose(m.p.y.value, y)
    x2, y2 = rng.randn(3, 2), rng.randn(4, 3)
    m.x = x2; m.p.y = y2
    assert np.allclose(m.x.value, x2) and np.allclose(m.p.y.value, y2) and isinstance(m.p.y, hb.param.Data)
    with pytest.raises(ValueError):
        m.p.y = rng.randn(4, 4)
    assert m.p.y._dtype == torch.float32 and m.p.z._dtype == torch.int32      # test_data.py:42-45


# ---------------------------------------------------------------------------------------------- test_model.py:137-147
def test_model_passes_keywords_to_setup():
    class ModelWithKeyword(hb.model.Model):
        def setUp(self, key1, key2):
            self.key1 = key1
            self.key2 = key2

    model = ModelWithKeyword(key1='hoge', key2='foo')
    assert model.key1 == 'hoge' and model.key2 == 'foo'


# ---------------------------------------------------------------------------------------------- test_model.py:76-105
def test_save_and_restore_by_long_name(tmp_path):
    class SquareModel2(hb.model.Model):
        def setUp(self):
            self.p = hb.param.Variable([2, 3], collections=['global1', 'global2'])
            self.q = hb.param.Variable([2, 3], collections=['global2'])
            self.v = hb.variationals.Gaussian([2, 3], collections=['global2'])

    m = SquareModel2()
    m.q = np.ones((2, 3))
    path = str(tmp_path / 'saved_file.dat')
    m.save(path)
    m2 = SquareModel2()
    m2.restore(path)
    assert np.allclose(m2.q.value, 1.0) and not np.allclose(m2.p.value, 1.0)
    m2.initialize()
    assert np.allclose(m2.q.value, 1.0)                   # initialize() does not undo a restore
    assert np.allclose(m2.p.value, m.p.value)
    # a sub-tree saves and restores on its own (test_model.py:91-105)
    m.v.q_mu = np.ones(6)
    path_v = str(tmp_path / 'saved_v.dat')
    m.v.save(path_v)
    m3 = SquareModel2()
    m3.v.restore(path_v)
    assert np.allclose(m3.v.q_mu.value, 1.0) and not np.allclose(m3.p.value, 1.0)
    m3.initialize()
    assert np.allclose(m3.v.q_mu.value, 1.0)


# ---------------------------------------------------------------------------------------------- test_variationals.py (host side)
def test_variational_children_are_variables_even_in_tf_mode():            # test_variationals.py:56-67
    for shape in ('fullrank', 'diagonal'):
        m = hb.model.Model()
        m.m = hb.variationals.Normal(10, n_layers=[3], q_shape=shape)
        with m.tf_mode():
            variables = m.get_variables()
        v = object.__getattribute__(m, 'm')
        assert any(x is v.q_mu for x in variables) and any(x is v.q_sqrt for x in variables)
        assert v._parent is m


def test_feeding_a_global_variational_does_nothing():                     # test_variationals.py:124-129
    m = hb.model.Model()
    m.m = hb.variationals.Normal(10, n_layers=[3])
    ref = m.m._tensor
    m.m.feed(np.ones(10))
    assert m.m._tensor is ref


def test_local_feed_size_is_the_same_inside_tf_mode():                    # test_variationals.py:166-180
    for shape, want in (('diagonal', 20), ('fullrank', 110)):
        m = hb.model.Model()
        m.m = hb.variationals.Normal([10], n_layers=[3], q_shape=shape, collections=graph_key.LOCAL)
        m.q = hb.variationals.Normal([10], n_layers=[3], n_batch=2, q_shape=shape)
        with m.tf_mode():
            inside = m.feed_size
            local = m.get_variables(graph_key.LOCAL)
        v = object.__getattribute__(m, 'm')
        assert inside == m.feed_size == want
        assert any(x is v.q_mu for x in local) and any(x is v.q_sqrt for x in local)
        assert m.q.q_mu._host.shape == (3, 2, 10)                          # storage: n_layers + [n_batch] + shape


def test_assigning_a_tensor_outside_tf_mode_replaces_the_attribute():     # test_variationals.py:224-234
    m = hb.model.Model()
    m.m = hb.variationals.Normal([10], n_layers=[3], collections=graph_key.LOCAL)
    m.m = torch.zeros(3, 2, 20)
    assert isinstance(m.m, torch.Tensor)


def test_variational_model_tree():                                        # test_variationals.py:236-264
    class VariationalModel(hb.model.Model):
        def setUp(self):
            self.q_global = hb.variationals.Normal(shape=[3])
            self.q_local = hb.variationals.Normal(shape=[3], collections=graph_key.LOCAL)
            self.x = hb.param.Variable(shape=[10, 6])

    m = VariationalModel()
    assert len(m.sorted_variables) == 3
    assert m.feed_size == 6 and m.q_local.is_local and not m.q_global.is_local


def test_scaled_families_accept_every_initialisation():                   # test_variationals.py:288-322 (construction part)
    for cls in (hb.variationals.Gaussian, hb.variationals.Beta):
        for mean, stddev in ((1.0, 0.5), (-1.0, 0.5), (0.0, 1.0)):
            g = cls(shape=[3, 2], n_layers=[1, 2], n_batch=0, mean=mean, stddev=stddev, scale_shape=[3, 2], scale_n_layers=[1, 2])
            assert g.q_mu._host.shape == (1, 2, 0, 6)                     # an EMPTY batch axis is legal upstream
            assert g.size == 6
    b = hb.variationals.Beta(shape=[3, 2], n_layers=[3], n_batch=2)       # test_variationals.py:349-356
    assert b.alpha._host.shape == (1, 2, 1, 1) and b.q_mu._host.shape == (3, 2, 6)


def test_sample_view_names_the_batch_axis():
    """Variational.tensor() views the flat sample as n_layers + [batch] + shape (variationals.py:112-119); an EMPTY batch
    axis (n_batch = 0 in test_variationals.py:288-322) must survive the view.  The sample itself is a CUDA kernel: here the
    drawn tensor is put in place by hand."""
    def view(v, t, S):
        v._tensor = t; v.transformed_tensor = t; v._S = S
        return tuple(v.tensor().shape)

    loc = hb.variationals.Normal([3, 2], n_layers=[2], collections=graph_key.LOCAL)
    assert view(loc, torch.zeros(2, 7, 6), 1) == (2, 7, 3, 2)
    assert view(loc, torch.zeros(4, 2, 7, 6), 4) == (4, 2, 7, 3, 2)                 # S samples: leading sample axis
    empty = hb.variationals.Normal([3, 2], n_layers=[1, 2], n_batch=0)
    assert view(empty, torch.zeros(1, 2, 0, 6), 1) == (1, 2, 0, 3, 2)
    plain = hb.variationals.Normal([3, 2], n_layers=[2])
    assert view(plain, torch.zeros(2, 6), 1) == (2, 3, 2)
