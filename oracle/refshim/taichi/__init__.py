"""A minimal stand-in for ``taichi==1.4.1`` -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Purpose: execute the UNMODIFIED text of the reference scripts (``/root/reference/2dvof.py``,
``3dvof.py``, ``test/forward_fct.py``) in this container, where the real Taichi cannot be installed
(py3.12, offline, requirements.txt:3), so that the hand-written oracles (``oracle/vof*_oracle.*``) and the
CUDA path can be pinned to the reference's own source instead of to a reading of it.  Driven by
``oracle/run_reference.py``; the product package never imports this.

What it models (the subset of the DSL the three scripts use), and how:

* ``ti.field(float, shape)`` is a zero-initialised fp32 NumPy array; a read inside a kernel returns
  ``np.float32``, a read in Python scope returns a Python float (so ``dx = x[3] - x[2]`` is a double that is
  the difference of two fp32 values, 2dvof.py:48).
* fp32 arithmetic is ``np.float32`` scalar arithmetic: IEEE, one rounding per operation, no contraction.
  NumPy >= 2 promotion (NEP 50) gives exactly Taichi's typing of constants: Python scalar (op) Python scalar
  is folded in double at "compile time"; Python scalar (op) fp32 value casts the scalar to fp32 first.
* loop indices are ``I32`` (an ``int`` subclass): i32 (op) Python float -> fp32, as in ``(i - imin) * dx``
  (2dvof.py:105).
* ``@ti.kernel`` / ``@ti.func`` re-compile the function through a small AST pass, as Taichi itself does.
  The pass changes nothing in the text; it only supplies the three DSL rules plain Python lacks:
  (1) a name is a kernel local only from its first assignment on, before that it is the module global
      (``F = var(0.0, 1.0, F[I])``, 2dvof.py:201, reads the global field and defines a local);
  (2) assigning a Python scalar creates a typed local (float -> fp32, int -> i32), so ``r = Lx / 12`` is an
      fp32 value and ``Ly - 3 * r`` is fp32 arithmetic (2dvof.py:155-156); re-assignment keeps the type;
  (3) ``a ** b``: folded in double for two Python scalars, repeated multiplication for a small integer
      exponent on a run-time value (Taichi's algebraic simplification), ``pow`` otherwise.
* each top-level ``for`` runs sequentially, cell by cell.  That is a valid serialisation of Taichi's parallel
  loop: within one loop of these scripts no cell reads what another cell writes, except the face arrays
  ``ax/ay`` that two cells write with the same value (2dvof.py:347-350).
* ``ti.max/min`` are n-ary through a right fold of the binary op (taichi/lang/ops.py); the binary op
  returns one of its operands.  Which operand is returned for equal values (sign of zero) is
  code-generator dependent in Taichi; ``TI_SHIM_MAXMIN=second`` flips the choice so that
  ``run_reference.py`` can show the fields do not depend on it.
* ``ti.GUI`` is headless: ``running`` turns False after a step budget and calls a hook after every step.

Not modelled: Taichi's default ``fast_math=True`` (its LLVM back end may contract/re-associate; this shim
is the IEEE evaluation of the source), autodiff, sparse SNodes, any other backend behaviour.
"""
from __future__ import annotations

import ast
import builtins
import inspect
import math
import os
import textwrap

import numpy as np

f32 = "f32"
i32 = "i32"
cpu = "cpu"
gpu = "gpu"

_F = np.float32
_depth = 0                       # > 0 while a kernel is executing
_hooks = {"after_kernel": None}  # run_reference.py installs a callback here
_MAXMIN_SECOND = os.environ.get("TI_SHIM_MAXMIN", "first") == "second"


def init(**_kw):
    return None


def template():
    return "template"


# ----------------------------------------------------------------------------- scalars
class I32(int):
    """i32 value.  int (op) int stays i32; i32 (op) Python float is an fp32 operation."""
    __slots__ = ()
    __array_ufunc__ = None          # np.float32 (op) I32 defers to the reflected method below

    def _f(self):
        return _F(int(self))

    def __add__(self, o):
        if type(o) is float or type(o) is _F:
            return self._f() + o
        r = int.__add__(self, o)
        return I32(r) if r is not NotImplemented else r

    def __radd__(self, o):
        if type(o) is float or type(o) is _F:
            return o + self._f()
        r = int.__radd__(self, o)
        return I32(r) if r is not NotImplemented else r

    def __sub__(self, o):
        if type(o) is float or type(o) is _F:
            return self._f() - o
        r = int.__sub__(self, o)
        return I32(r) if r is not NotImplemented else r

    def __rsub__(self, o):
        if type(o) is float or type(o) is _F:
            return o - self._f()
        r = int.__rsub__(self, o)
        return I32(r) if r is not NotImplemented else r

    def __mul__(self, o):
        if type(o) is float or type(o) is _F:
            return self._f() * o
        r = int.__mul__(self, o)
        return I32(r) if r is not NotImplemented else r

    def __rmul__(self, o):
        if type(o) is float or type(o) is _F:
            return o * self._f()
        r = int.__rmul__(self, o)
        return I32(r) if r is not NotImplemented else r

    def __truediv__(self, o):       # Taichi: i32 / x is a floating-point division in default_fp
        return self._f() / (_F(int(o)) if isinstance(o, int) else o)

    def __rtruediv__(self, o):
        return (_F(int(o)) if isinstance(o, int) else o) / self._f()

    def __floordiv__(self, o):
        r = int.__floordiv__(self, o)
        return I32(r) if r is not NotImplemented else r

    def __mod__(self, o):
        r = int.__mod__(self, o)
        return I32(r) if r is not NotImplemented else r

    def __neg__(self):
        return I32(int.__neg__(self))


def _typed(v):
    """Rule (2): what a Python scalar becomes when it is stored in a kernel local / passed to a func."""
    t = type(v)
    if t is float:
        return _F(v)
    if t is int:
        return I32(v)
    if t is np.float64:                                 # a double must never leak into kernel arithmetic
        raise TypeError("float64 value inside a kernel: the shim mis-typed an expression")
    return v


def _recast(old, v):
    """Store to an existing local keeps the local's type."""
    t = type(old)
    if t is _F:
        return v if type(v) is _F else _F(v)
    if t is I32:
        return v if type(v) is I32 else I32(int(v))
    return _typed(v)


def _pow(a, b):
    """Rule (3)."""
    if type(a) in (int, float) and type(b) in (int, float):
        return a ** b                                    # compile-time constant, double
    if isinstance(b, int) and 0 < int(b) <= 32:          # a ** n -> a * a * ... (left to right)
        if type(a) is float:
            a = _F(a)
        r = a
        for _ in range(int(b) - 1):
            r = r * a
        return r
    return _F(a) ** _F(b)


class IVec(tuple):
    """Index vector of ``ti.grouped`` (2dvof.py:199, 460-482)."""
    __slots__ = ()

    def __floordiv__(self, o):
        return IVec(I32(int(c) // int(o)) for c in self)


class Vector:
    """``ti.Vector([...])`` as an fp32 array; ``ti.Vector.field`` as a field with a trailing axis."""

    def __new__(cls, items):
        return np.array([_F(c) for c in items], dtype=_F)

    @staticmethod
    def field(n, dtype=float, shape=()):
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        return Field(shape + (n,), vector=True)


# ------------------------------------------------------------------------------ fields
class Field:
    def __init__(self, shape, vector=False):
        shape = (shape,) if isinstance(shape, int) else tuple(shape)
        self.a = np.zeros(shape, dtype=_F)
        self.vector = vector
        self.shape = shape[:-1] if vector else shape

    def __getitem__(self, idx):
        if idx is None:
            idx = ()
        v = self.a[idx]
        if _depth:
            return v
        return float(v) if np.ndim(v) == 0 else v

    def __setitem__(self, idx, val):
        if idx is None:
            idx = ()
        self.a[idx] = val

    def __iter__(self):                                # struct-for: ``for i, j in F`` / ``for i in xm``
        if len(self.shape) == 1:
            return (I32(i) for i in range(self.shape[0]))
        return (tuple(I32(c) for c in idx) for idx in np.ndindex(*self.shape))

    def from_numpy(self, arr):
        self.a[...] = np.asarray(arr).astype(_F)        # Taichi casts to the field dtype

    def to_numpy(self):
        return self.a.copy()

    def fill(self, val):
        self.a[...] = val


def field(dtype=float, shape=(), **_kw):
    return Field(shape)


def grouped(f):
    return (IVec(I32(c) for c in idx) for idx in np.ndindex(*f.shape))


def ndrange(*args):
    rs = []
    for a in args:
        lo, hi = (0, a) if isinstance(a, int) else a
        rs.append(range(int(lo), int(hi)))
    if len(rs) == 1:
        return (I32(i) for i in rs[0])
    return _nd(rs)


def _nd(rs):
    if len(rs) == 2:
        r1 = [I32(j) for j in rs[1]]
        for i in rs[0]:
            i = I32(i)
            for j in r1:
                yield i, j
    elif len(rs) == 3:
        r1 = [I32(j) for j in rs[1]]
        r2 = [I32(k) for k in rs[2]]
        for i in rs[0]:
            i = I32(i)
            for j in r1:
                for k in r2:
                    yield i, j, k
    else:
        raise NotImplementedError(len(rs))


# -------------------------------------------------------------------------------- ops
def _is_py(v):
    return type(v) in (int, float, bool)


def _max2(a, b):
    if _is_py(a) and _is_py(b):
        return builtins.max(a, b)
    if _is_py(a):
        a = _typed(a)
    if _is_py(b):
        b = _typed(b)
    if isinstance(a, int) and isinstance(b, int):
        return I32(builtins.max(int(a), int(b)))
    a, b = _F(a), _F(b)
    if _MAXMIN_SECOND:
        return a if a > b else b
    return b if b > a else a


def _min2(a, b):
    if _is_py(a) and _is_py(b):
        return builtins.min(a, b)
    if _is_py(a):
        a = _typed(a)
    if _is_py(b):
        b = _typed(b)
    if isinstance(a, int) and isinstance(b, int):
        return I32(builtins.min(int(a), int(b)))
    a, b = _F(a), _F(b)
    if _MAXMIN_SECOND:
        return a if a < b else b
    return b if b < a else a


def max(*args):                                         # noqa: A001  (taichi/lang/ops.py: right fold)
    if len(args) == 2:
        return _max2(args[0], args[1])
    return _max2(args[0], max(*args[1:]))


def min(*args):                                         # noqa: A001
    if len(args) == 2:
        return _min2(args[0], args[1])
    return _min2(args[0], min(*args[1:]))


def _unary(np_fn, py_fn):
    def op(x):
        if _is_py(x):
            return py_fn(x)                            # Python scalar: folded in double
        return np_fn(_F(x))
    return op


sqrt = _unary(np.sqrt, math.sqrt)
sin = _unary(np.sin, math.sin)
cos = _unary(np.cos, math.cos)
abs = _unary(np.abs, builtins.abs)                      # noqa: A001


# --------------------------------------------------------------- kernel / func decorators
class _Scope(ast.NodeTransformer):
    """The three DSL rules, applied in source order (see the module docstring)."""

    def __init__(self, argnames):
        self.local = set(argnames)

    @staticmethod
    def _l(name):
        return "_l_" + name

    def visit_Name(self, node):
        if node.id in self.local:
            return ast.copy_location(ast.Name(self._l(node.id), node.ctx), node)
        return node

    def visit_arg(self, node):
        node.arg = self._l(node.arg)
        node.annotation = None
        return node

    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Pow):
            call = ast.Call(ast.Name("__ti_pow__", ast.Load()), [node.left, node.right], [])
            return ast.copy_location(call, node)
        return node

    def _store(self, target, value):
        """value expression (already visited) -> list of statements storing it into target."""
        if isinstance(target, ast.Name):
            name = target.id
            if name in self.local:
                fn, args = "__ti_recast__", [ast.Name(self._l(name), ast.Load()), value]
            else:
                fn, args = "__ti_typed__", [value]
                self.local.add(name)
            return ast.Assign([ast.Name(self._l(name), ast.Store())],
                              ast.Call(ast.Name(fn, ast.Load()), args, []))
        for t in ast.walk(target):                       # ``a, b = f()``: names become locals
            if isinstance(t, ast.Name) and isinstance(t.ctx, ast.Store):
                self.local.add(t.id)
        return ast.Assign([self.visit(target)], value)

    def visit_Assign(self, node):
        value = self.visit(node.value)
        out = []
        for tgt in node.targets:
            if isinstance(tgt, ast.Tuple) and isinstance(value, ast.Tuple) and len(tgt.elts) == len(value.elts):
                for t, v in zip(tgt.elts, value.elts):   # ``cx, cy = Lx / 2, 2 * r``
                    out.append(self._store(t, v))
            else:
                out.append(self._store(tgt, value))
        return [ast.copy_location(s, node) for s in out]

    def visit_AugAssign(self, node):
        if isinstance(node.target, ast.Name):
            load = ast.Name(node.target.id, ast.Load())
            return self.visit_Assign(ast.copy_location(
                ast.Assign([node.target], ast.BinOp(load, node.op, node.value)), node))
        self.generic_visit(node)
        return node

    def visit_For(self, node):
        node.iter = self.visit(node.iter)
        outer = set(self.local)
        for t in ast.walk(node.target):
            if isinstance(t, ast.Name):
                self.local.add(t.id)
        node.target = self.visit(node.target)
        node.body = self._block(node.body)
        self.local = outer
        return node

    def _block(self, body):
        """Locals defined inside a block end with it (Taichi scopes them; 2dvof.py:150-155 defines ``r``
        in two sibling branches)."""
        outer = set(self.local)
        out = [s for b in body for s in self._stmts(b)]
        self.local = outer
        return out

    def _stmts(self, stmt):
        r = self.visit(stmt)
        return r if isinstance(r, list) else [r]

    def visit_If(self, node):
        node.test = self.visit(node.test)
        node.body = self._block(node.body)
        node.orelse = self._block(node.orelse)
        return node


def _compile(fn):
    src = textwrap.dedent(inspect.getsource(fn))
    tree = ast.parse(src)
    fdef = tree.body[0]
    assert isinstance(fdef, ast.FunctionDef)
    fdef.decorator_list = []
    anns = [a.annotation for a in fdef.args.args]
    argnames = [a.arg for a in fdef.args.args]
    sc = _Scope(argnames)
    fdef.args = sc.visit(fdef.args)
    fdef.body = [s for b in fdef.body for s in sc._stmts(b)]
    ast.fix_missing_locations(tree)
    ast.increment_lineno(tree, fn.__code__.co_firstlineno - 1)
    g = fn.__globals__
    g.setdefault("__ti_pow__", _pow)
    g.setdefault("__ti_typed__", _typed)
    g.setdefault("__ti_recast__", _recast)
    ns = {}
    exec(compile(tree, fn.__code__.co_filename, "exec"), g, ns)
    return ns[fdef.name], anns, g


def func(fn):
    body, _anns, _g = _compile(fn)

    def call(*args):
        return body(*[_typed(a) for a in args])        # by value; Python scalars become typed locals
    call.__name__ = fn.__name__
    return call


def kernel(fn):
    body, anns, g = _compile(fn)
    kinds = []
    for a in anns:
        try:
            kinds.append(eval(compile(ast.Expression(a), "<ann>", "eval"), g) if a is not None else None)
        except Exception:
            kinds.append(None)

    def launch(*args):
        global _depth
        conv = []
        for a, k in zip(args, kinds):
            if k == i32:
                a = I32(int(a))
            elif k == f32 or k is float:
                a = _F(a)
            conv.append(a)
        _depth += 1
        try:
            body(*conv)
        finally:
            _depth -= 1
        cb = _hooks["after_kernel"]
        if cb is not None:
            cb(fn.__name__)
    launch.__name__ = fn.__name__
    return launch


# --------------------------------------------------------------------------------- GUI
class GUI:
    """Headless ``ti.GUI``: ``running`` is the step budget of ``while gui.running`` (2dvof.py:505)."""
    RELEASE = "release"
    SPACE = "space"
    budget = 0           # steps to run; set by run_reference.py
    on_step = None       # callback(step_index) after every completed step

    def __init__(self, *a, **kw):
        self._polls = 0
        self._stop = False

    @property
    def running(self):
        done = self._polls
        self._polls += 1
        if done and GUI.on_step is not None:
            GUI.on_step(done)
        return (not self._stop) and done < GUI.budget

    @running.setter
    def running(self, v):
        self._stop = not v

    def get_events(self, *a):
        return []

    def set_image(self, *a, **kw):
        pass

    def show(self, *a, **kw):
        pass

    def contour(self, *a, **kw):
        pass
