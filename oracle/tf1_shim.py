"""A minimal stand-in for the TensorFlow-1.x API surface that the reference's graph-definition
code calls -- TEST INFRASTRUCTURE (used only by oracle/run_reference.py to execute the reference's
own, unmodified source in this container and write tests/golden/ref_*.npz).

TensorFlow (``tensorflow>=1.0``, setup.py:40, unpinned) is absent from /root/reference and from
this image.  What the reference needs from it on the hot path is ordinary dense arithmetic with
reverse-mode autodiff: matmul, add_n, elementwise ops, reductions, stack/concat/slicing,
``function.Defun`` with a custom ``grad_func``, ``tf.train.AdamOptimizer`` and ``Session.run``.
This module restates exactly those published semantics on top of torch-CPU:

  * graph mode: every ``tf.*`` call builds a lazy ``Node``; ``Session.run(fetches, feed_dict)``
    evaluates the requested nodes once (memoised per run) with the current variable values;
  * ``Optimizer.compute_gradients(loss)`` -> nodes evaluating d loss / d variable by autograd;
  * ``function.Defun(..., grad_func=g)``: forward = the python body, backward = ``g(*inputs, grad)``
    (TF calls grad_func with the op inputs followed by the output gradients);
  * ``tf.train.AdamOptimizer``: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v moments; theta -= lr_t*m/(sqrt(v)+eps)
    with beta1=0.9, beta2=0.999, epsilon=1e-8 (TF-1 documentation);
  * ``tf.nn.l2_loss(x) = sum(x**2)/2``.

``set_float(torch.float32 | torch.float64)`` chooses what ``tf.float32`` maps to: float32 is the
reference's real dtype; float64 yields the "reference semantics at fp64" goldens that the fp64
CUDA path is compared against at tight tolerance.
"""
import contextlib
import sys
import types

import numpy as np
import torch

_FLOAT = torch.float32
_DEFAULT_SESSION = None
_ALL_VARIABLES = []


def set_float(dtype):
    global _FLOAT
    _FLOAT = dtype


class _DType:
    def __init__(self, name):
        self.name = name

    def resolve(self):
        return {'float32': _FLOAT, 'float64': torch.float64, 'complex64': torch.complex64,
                'int32': torch.int32}[self.name]


float32, float64, complex64, int32 = _DType('float32'), _DType('float64'), _DType('complex64'), _DType('int32')


def _td(dtype):
    if dtype is None:
        return _FLOAT
    if isinstance(dtype, _DType):
        return dtype.resolve()
    return dtype


class _Ctx:
    """One Session.run: memo of evaluated nodes, leaf tensors of variables, feeds."""

    def __init__(self, feeds=None):
        self.cache = {}
        self.leaves = {}
        self.feeds = feeds or {}

    def leaf(self, var):
        if var not in self.leaves:
            self.leaves[var] = var.value.detach().clone().requires_grad_(True)
        return self.leaves[var]


class Node:
    __array_ufunc__ = None          # make numpy scalars defer to __rmul__/__radd__ ...

    def __init__(self, fn, name=None):
        self._fn = fn
        self.name = name

    def _eval(self, ctx):
        k = id(self)
        if k not in ctx.cache:
            ctx.cache[k] = self._fn(ctx)
        return ctx.cache[k]

    # tf.Tensor.eval()
    def eval(self, feed_dict=None, session=None):
        sess = session or _DEFAULT_SESSION
        return sess.run(self, feed_dict=feed_dict)

    def __getitem__(self, idx):
        return Node(lambda ctx: self._eval(ctx)[idx])

    def __add__(self, o): return _bin(self, o, lambda a, b: a + b)
    def __radd__(self, o): return _bin(o, self, lambda a, b: a + b)
    def __sub__(self, o): return _bin(self, o, lambda a, b: a - b)
    def __rsub__(self, o): return _bin(o, self, lambda a, b: a - b)
    def __mul__(self, o): return _bin(self, o, lambda a, b: a * b)
    def __rmul__(self, o): return _bin(o, self, lambda a, b: a * b)
    def __truediv__(self, o): return _bin(self, o, lambda a, b: a / b)
    def __rtruediv__(self, o): return _bin(o, self, lambda a, b: a / b)
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __neg__(self): return Node(lambda ctx: -self._eval(ctx))


def _val(x, ctx):
    if isinstance(x, Node):
        return x._eval(ctx)
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (list, tuple)) and any(isinstance(e, Node) for e in x):
        return torch.stack([_val(e, ctx) for e in x])
    if isinstance(x, (int, float, np.floating, np.integer)):
        return x                       # python scalars stay weakly typed, like TF constants-from-python
    return torch.as_tensor(np.asarray(x), dtype=_FLOAT)


def _bin(a, b, op):
    return Node(lambda ctx: op(_val(a, ctx), _val(b, ctx)))


def _un(a, op):
    return Node(lambda ctx: op(_val(a, ctx)))


class Variable(Node):
    def __init__(self, initial_value, trainable=True, dtype=None, name=None):
        self.trainable = trainable
        self.name = name
        self._init = initial_value
        self._dtype = dtype
        self.value = None
        Node.__init__(self, None, name)
        _ALL_VARIABLES.append(self)

    def initialize(self):
        v = _val(self._init, _Ctx())
        if not isinstance(v, torch.Tensor):
            v = torch.as_tensor(v)
        self.value = v.detach().clone().to(_td(self._dtype))

    def _eval(self, ctx):
        if self.value is None:          # shapes are sometimes needed at graph-construction time (tf.unstack)
            self.initialize()
        return ctx.leaf(self) if self.trainable else self.value

    def assign(self, value):
        def do(ctx):
            self.value = torch.as_tensor(np.asarray(value), dtype=self.value.dtype).clone()
            return self.value
        return Node(do)


# ---- ops -----------------------------------------------------------------------------------------
def constant(value, dtype=None, name=None, shape=None):
    t = torch.as_tensor(np.asarray(value)).to(_td(dtype))
    return Node(lambda ctx: t, name)


def zeros(shape, dtype=None, name=None):
    return Node(lambda ctx: torch.zeros(_shape(shape, ctx), dtype=_td(dtype)), name)


def ones(shape, dtype=None, name=None):
    return Node(lambda ctx: torch.ones(_shape(shape, ctx), dtype=_td(dtype)), name)


def _shape(s, ctx):
    if isinstance(s, Node):
        s = s._eval(ctx)
    return tuple(int(x) for x in s)


def shape(x):
    return Node(lambda ctx: tuple(_val(x, ctx).shape))


def placeholder(dtype, shape=None, name=None):
    node = Node(None, name)
    node._fn = lambda ctx: torch.as_tensor(ctx.feeds[node], dtype=_td(dtype))
    return node


def matmul(a, b, a_is_sparse=False, b_is_sparse=False, name=None, transpose_a=False, transpose_b=False):
    return _bin(a, b, lambda x, y: torch.matmul(x, y))       # sparsity flags are hints only in TF


def add_n(inputs, name=None):
    def f(ctx):
        vals = [_val(i, ctx) for i in inputs]
        out = vals[0]
        for v in vals[1:]:
            out = out + v
        return out
    return Node(f)


def multiply(a, b, name=None): return _bin(a, b, lambda x, y: x * y)
def add(a, b, name=None): return _bin(a, b, lambda x, y: x + y)
def subtract(a, b, name=None): return _bin(a, b, lambda x, y: x - y)
def square(a, name=None): return _un(a, lambda x: x * x)
def sin(a, name=None): return _un(a, torch.sin)
def transpose(a, name=None): return _un(a, lambda x: x.t() if x.dim() == 2 else x.permute(*reversed(range(x.dim()))))
def cast(a, dtype=None): return _un(a, lambda x: x.to(_td(dtype)))


def reduce_sum(a, axis=None, name=None):
    return _un(a, lambda x: torch.sum(x) if axis is None else torch.sum(x, dim=axis))


def stack(values, axis=0, name=None):
    return Node(lambda ctx: torch.stack([torch.as_tensor(_val(v, ctx), dtype=_FLOAT) if not isinstance(_val(v, ctx), torch.Tensor)
                                         else _val(v, ctx) for v in values], dim=axis), name)


def unstack(value, axis=0, num=None, name=None):
    # the number of outputs must be known at graph-construction time: evaluate the shape eagerly
    n = _val(value, _Ctx()).shape[axis] if num is None else num
    return [Node((lambda i: lambda ctx: torch.select(_val(value, ctx), axis, i))(i)) for i in range(n)]


def concat(values, axis, name=None):
    return Node(lambda ctx: torch.cat([_val(v, ctx) for v in values], dim=axis))


def tile(a, multiples, name=None): return _un(a, lambda x: x.repeat(*multiples))
def reshape(a, shp, name=None): return _un(a, lambda x: x.reshape(*[int(s) for s in shp]))


@contextlib.contextmanager
def name_scope(name):
    yield


@contextlib.contextmanager
def device(name):
    yield


class _NN:
    @staticmethod
    def l2_loss(x, name=None):
        return _un(x, lambda v: torch.sum(v * v) / 2)


nn = _NN()


class Graph:
    @contextlib.contextmanager
    def as_default(self):
        yield self


class ConfigProto:
    def __init__(self, **kw):
        pass


class _Init:
    def run(self, feed_dict=None, session=None):
        for v in _ALL_VARIABLES:
            if v.value is None:
                v.initialize()


def global_variables_initializer():
    return _Init()


class Session:
    def __init__(self, graph=None, config=None):
        self.graph = graph

    def __enter__(self):
        global _DEFAULT_SESSION
        _DEFAULT_SESSION = self
        return self

    def __exit__(self, *a):
        return False

    def run(self, fetches, feed_dict=None):
        ctx = _Ctx(feed_dict)
        single = not isinstance(fetches, (list, tuple))
        out = []
        for f in ([fetches] if single else fetches):
            v = f._eval(ctx)
            if isinstance(v, torch.Tensor):
                v = v.detach().numpy().copy()
                if v.ndim == 0:
                    v = v[()]
            out.append(v)
        return out[0] if single else out


class _AdamOptimizer:
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.t = 0
        self.slots = {}

    def compute_gradients(self, loss, var_list=None):
        vs = var_list or [v for v in _ALL_VARIABLES if v.trainable]
        out = []
        for v in vs:
            def g(ctx, v=v):
                L = loss._eval(ctx)
                (gr,) = torch.autograd.grad(L, ctx.leaf(v), retain_graph=True)
                return gr
            out.append((Node(g), v))
        return out

    def apply_gradients(self, grads_and_vars):
        def f(ctx):
            lr = _val(self.lr, ctx)
            self.t += 1
            dt = _FLOAT
            b1, b2 = torch.tensor(self.b1, dtype=dt), torch.tensor(self.b2, dtype=dt)
            lr_t = torch.as_tensor(lr, dtype=dt) * torch.sqrt(1 - b2 ** self.t) / (1 - b1 ** self.t)
            for gnode, var in grads_and_vars:
                g = gnode._eval(ctx)
                m, v = self.slots.setdefault(var, (torch.zeros_like(var.value), torch.zeros_like(var.value)))
                m = b1 * m + (1 - b1) * g
                v = b2 * v + (1 - b2) * g * g
                self.slots[var] = (m, v)
                var.value = (var.value - lr_t * m / (torch.sqrt(v) + self.eps)).detach()
            return None
        return Node(f)


class _Saver:
    def __init__(self, *a, **k):
        pass


class _Train:
    AdamOptimizer = _AdamOptimizer
    Saver = _Saver


train = _Train()


# ---- tensorflow.python.framework.function ----------------------------------------------------------
class _DefunOp:
    def __init__(self, pyfunc, grad_func):
        self.pyfunc, self.grad_func = pyfunc, grad_func

    def _run_body(self, vals):
        out = self.pyfunc(*[Node((lambda v: lambda ctx: v)(v)) for v in vals])
        ctx = _Ctx()
        if isinstance(out, (list, tuple)):
            return [_val(o, ctx) for o in out]
        return _val(out, ctx)

    def _apply(self, vals):
        if self.grad_func is None:
            return self._run_body(vals)
        op = self

        class Fn(torch.autograd.Function):
            @staticmethod
            def forward(fctx, *inputs):
                fctx.save_for_backward(*inputs)
                with torch.no_grad():
                    return op._run_body([i.detach() for i in inputs])

            @staticmethod
            def backward(fctx, grad):
                inputs = [i.detach() for i in fctx.saved_tensors]
                with torch.no_grad():
                    grads = op.grad_func._run_body(inputs + [grad])
                return tuple(grads)

        return Fn.apply(*vals)

    def __call__(self, *args):
        return Node(lambda ctx: self._apply([_val(a, ctx) for a in args]))


class Defun:
    def __init__(self, *input_types, **kwargs):
        self.grad_func = kwargs.get('grad_func')

    def __call__(self, pyfunc):
        return _DefunOp(pyfunc, self.grad_func)


def install():
    """Register this module as ``tensorflow`` (+ the two sub-modules the reference imports)."""
    me = sys.modules[__name__]
    sys.modules['tensorflow'] = me
    py = types.ModuleType('tensorflow.python')
    fw = types.ModuleType('tensorflow.python.framework')
    fn = types.ModuleType('tensorflow.python.framework.function')
    fn.Defun = Defun
    ops_mod = types.ModuleType('tensorflow.python.framework.ops')
    fw.function, fw.ops = fn, ops_mod
    py.framework = fw
    me.python = py
    sys.modules['tensorflow.python'] = py
    sys.modules['tensorflow.python.framework'] = fw
    sys.modules['tensorflow.python.framework.function'] = fn
    sys.modules['tensorflow.python.framework.ops'] = ops_mod


def reset():
    """Forget all variables / sessions (one Grape() call = one fresh graph)."""
    global _DEFAULT_SESSION
    del _ALL_VARIABLES[:]
    _DEFAULT_SESSION = None
