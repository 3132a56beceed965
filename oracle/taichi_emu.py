"""A stand-in for the `taichi` package that EXECUTES the reference's kernels in plain Python - TEST INFRASTRUCTURE
ONLY (used by tools/make_golden_ref.py in the build container, where /root/reference exists; never by the product).

Why: Taichi cannot be installed here, so the reference cannot run as written.  But its kernels are Python source
that uses a small, fixed set of Taichi constructs (inventory: ti.Vector / ti.Matrix / fields on pointer-dense SNodes,
ti.static, ti.cast, ti.round/floor/abs/min, struct-for over sparse fields, ti.atomic_add on counters).  This module
implements exactly those with Taichi's semantics:

  * default_fp = f32, default_ip = i32: vectors are float32 / int32 numpy arrays, field loads return np.float32 (f16
    fields are widened on load and rounded to binary16 on store, like Taichi's f16 fields), Python literals stay
    "weak" (NumPy 2 promotion), so `1.0 / (z * z)` is an f32 operation as in the compiled kernel;
  * sparse fields: reads of inactive cells return 0, writes activate, `deactivate_all()` clears, `for i, j, k in f`
    visits the active cells (first-touch order - one legal serial schedule of the racy parallel loops);
  * `ti.round` rounds half away from zero (llvm.round), integer casts truncate;
  * `ti.atomic_add(x, v)` cannot mutate a Python lvalue, so the loader rewrites `t = ti.atomic_add(X, v)` into
    `t = X; X = X + v` (and the bare call into `X = X + v`) in the AST before compiling - the only source
    transformation applied; `range` is shadowed by a version that truncates float bounds as Taichi does.

`load_reference(root)` imports the unmodified files of `<root>/taichi_slam/mapping/` through this stand-in and returns
the package.  What comes out is the reference's own code run serially in f32/f16 - not Taichi's LLVM code (no FMA
contraction, no races), which is the documented residual of the "parity unpinned" statement.
"""
import ast
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types

import numpy as np


# ---------------------------------------------------------------------------------------------------------------
# dtypes
# ---------------------------------------------------------------------------------------------------------------
class _DT:
    def __init__(self, name, np_dtype):
        self.name, self.np = name, np_dtype

    def __repr__(self):
        return "ti." + self.name


f16, f32, f64 = _DT("f16", np.float16), _DT("f32", np.float32), _DT("f64", np.float64)
i8, i16, i32, i64 = _DT("i8", np.int8), _DT("i16", np.int16), _DT("i32", np.int32), _DT("i64", np.int64)
u8, u16, u32 = _DT("u8", np.uint8), _DT("u16", np.uint16), _DT("u32", np.uint32)
int32, float32 = i32, f32


def _npdt(dt):
    if dt is None:
        return None
    if isinstance(dt, _DT):
        return dt.np
    if dt is float:
        return np.float32
    if dt is int:
        return np.int32
    return np.dtype(dt).type


# ---------------------------------------------------------------------------------------------------------------
# Taichi's value typing.  Three value classes matter on this path: f16 (values loaded from ti.f16 fields and whatever
# is computed from f16 operands only), f32 (default_fp: literals, ti.f32 fields, compile-time Python floats) and
# integers.  A binary operation yields the wider class (taichi `promoted_type`): f16 (+) f16 -> f16 (rounded to
# binary16 after the operation), f16 (+) f32 -> f32, int (+) float -> that float class, int / int -> f32.
# ---------------------------------------------------------------------------------------------------------------
def _cls(x):
    if isinstance(x, H16):
        return 1
    if isinstance(x, Vec):
        return 1 if x.a.dtype == np.float16 else (2 if x.a.dtype.kind == "f" else 0)
    if isinstance(x, np.ndarray):
        return 1 if x.dtype == np.float16 else (2 if x.dtype.kind == "f" else 0)
    if isinstance(x, (float, np.floating)):
        return 1 if isinstance(x, np.float16) else 2
    if isinstance(x, (list, tuple)):
        return max((_cls(v) for v in x), default=0)
    return 0


_NP = {0: np.int32, 1: np.float16, 2: np.float32}


def _raw(x, dt):
    if isinstance(x, H16):
        return dt(x.v)
    if isinstance(x, Vec):
        return x.a.astype(dt, copy=False)
    if isinstance(x, (list, tuple)):
        return np.array([float(v) if _cls(v) else int(v) for v in x]).astype(dt)
    if isinstance(x, np.ndarray):
        return x.astype(dt, copy=False)
    return dt(x)


def _wrap(r):
    if isinstance(r, np.ndarray) and r.ndim:
        return Vec(r)
    r = r[()] if isinstance(r, np.ndarray) else r
    if isinstance(r, np.float16):
        return H16(r)
    if isinstance(r, np.floating):
        return np.float32(r)
    if isinstance(r, (np.bool_, bool)):
        return bool(r)
    return int(r)


def _binop(a, b, op, truediv=False):
    c = max(_cls(a), _cls(b))
    if truediv and c == 0:
        c = 2
    dt = _NP[c]
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        return _wrap(op(_raw(a, dt), _raw(b, dt)))


class _Arith:
    def __add__(self, o): return _binop(self, o, np.add)
    def __radd__(self, o): return _binop(o, self, np.add)
    def __sub__(self, o): return _binop(self, o, np.subtract)
    def __rsub__(self, o): return _binop(o, self, np.subtract)
    def __mul__(self, o): return _binop(self, o, np.multiply)
    def __rmul__(self, o): return _binop(o, self, np.multiply)
    def __truediv__(self, o): return _binop(self, o, np.divide, True)
    def __rtruediv__(self, o): return _binop(o, self, np.divide, True)

    def __iadd__(self, o):
        """`field[idx] += v`: Taichi lowers it to an atomic add on the destination type - the addend is converted to the
        field's type FIRST, then added (f16 field: round(v) then an f16 addition)."""
        c = _cls(self)
        dt = _NP[c]
        return _wrap(np.add(_raw(self, dt), _raw(o, dt)))


class H16(_Arith):
    """A scalar of Taichi type f16."""
    __slots__ = ("v",)
    __array_priority__ = 100
    __array_ufunc__ = None   # numpy scalars defer to the reflected operators below

    def __init__(self, v):
        self.v = np.float16(v)

    def __float__(self): return float(self.v)
    def __int__(self): return int(self.v)
    def __repr__(self): return f"H16({float(self.v)!r})"
    def __neg__(self): return H16(-self.v)
    def __abs__(self): return H16(abs(self.v))
    def __lt__(self, o): return float(self) < float(o)
    def __le__(self, o): return float(self) <= float(o)
    def __gt__(self, o): return float(self) > float(o)
    def __ge__(self, o): return float(self) >= float(o)
    def __eq__(self, o): return float(self) == float(o)
    def __ne__(self, o): return float(self) != float(o)
    def __hash__(self): return hash(float(self.v))
    def __bool__(self): return bool(self.v)


class Vec(_Arith):
    """ti.Vector / ti.Matrix value (numpy array of f16 / f32 / i32 in `.a`)."""
    __array_priority__ = 100
    __array_ufunc__ = None

    def __init__(self, a):
        self.a = a

    def __len__(self): return self.a.shape[0]
    def __eq__(self, o): return [bool(x) for x in (self.a == (o.a if isinstance(o, Vec) else np.asarray(o))).reshape(-1)]  # elementwise, for all()/any() (dense_esdf.py:268)
    __hash__ = object.__hash__
    def __iter__(self): return (_wrap(v) for v in self.a) if self.a.ndim == 1 else (Vec(r) for r in self.a)
    def __repr__(self): return f"Vec({self.a!r})"
    def __neg__(self): return Vec(-self.a)
    def __array__(self, dtype=None, copy=None): return self.a if dtype is None else self.a.astype(dtype)

    def __getitem__(self, i):
        if isinstance(i, tuple):
            i = tuple(int(v) if not isinstance(v, slice) else v for v in i)
            if self.a.ndim == 2 and any(isinstance(v, slice) for v in i):
                r = self.a[i]
                return Vec(r.reshape(1, -1))  # taichi: m[r, :] is a 1 x n matrix
        r = self.a[i]
        return _wrap(r) if not (isinstance(r, np.ndarray) and r.ndim) else Vec(r)

    def __setitem__(self, i, v):
        self.a[i] = _raw(v, self.a.dtype.type)

    def __matmul__(self, o):
        c = max(_cls(self), _cls(o))
        dt = _NP[c if c else 2]
        A, B = _raw(self, dt), _raw(o, dt)
        if A.ndim == 2 and B.ndim == 1:  # row-by-row accumulation in the operand type, sum in index order
            out = np.zeros(A.shape[0], dt)
            for r in range(A.shape[0]):
                acc = dt(0)
                for k in range(A.shape[1]):
                    acc = dt(acc + dt(A[r, k] * B[k]))
                out[r] = acc
            return Vec(out)
        return Vec((A @ B).astype(dt))

    def dot(self, o):
        c = max(_cls(self), _cls(o))
        dt = _NP[c if c else 2]
        A, B = _raw(self, dt), _raw(o, dt)
        acc = dt(0)
        for k in range(A.shape[0]):
            acc = dt(acc + dt(A[k] * B[k]))
        return _wrap(acc)

    def norm(self, eps=0):
        dt = self.a.dtype.type if self.a.dtype.kind == "f" else np.float32
        s = _raw(self.dot(self), dt)
        return _wrap(np.sqrt(dt(s + dt(eps)))) if eps else _wrap(np.sqrt(s))

    def normalized(self, eps=0):
        return self / (self.norm() + eps) if eps else self / self.norm()

    def cross(self, o):
        a, b = self, o
        return _vec([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])

    def transpose(self, *a):
        if self.a.ndim == 2 and self.a.shape[0] == 1:
            return Vec(self.a.reshape(-1))   # (1 x n)^T is an n-vector
        return Vec(self.a.T) if self.a.ndim == 2 else self

    def to_numpy(self):
        return self.a.copy()


def _vec(data, dt=None):
    npd = _npdt(dt)
    if isinstance(data, Vec):
        data = data.a
    if npd is None:
        c = _cls(list(data) if not isinstance(data, np.ndarray) else data)
        npd = _NP[c]
        if isinstance(data, np.ndarray) and data.dtype.kind == "f" and data.dtype != np.float16:
            npd = np.float32
    if isinstance(data, np.ndarray):
        return Vec(data.astype(npd))
    flat = []

    def conv(x):
        if isinstance(x, (list, tuple, Vec)):
            return [conv(v) for v in x]
        return float(x) if _cls(x) else int(x)

    return Vec(np.array(conv(data)).astype(npd))


class _VectorMeta(type):
    def __call__(cls, data, dt=None, **k):
        return _vec(data, dt)


class Vector(metaclass=_VectorMeta):
    @staticmethod
    def field(n, dtype=None, shape=None, **k):
        return Field(dtype, shape=shape, elem=(n,))

    @staticmethod
    def zero(dt, n):
        return _vec(np.zeros(n), dt)


class Matrix(metaclass=_VectorMeta):
    @staticmethod
    def field(n, m, dtype=None, shape=None, **k):
        return Field(dtype, shape=shape, elem=(n, m))

    @staticmethod
    def identity(dt, n):
        return _vec(np.eye(n), dt)

    @staticmethod
    def zero(dt, n, m=None):
        return _vec(np.zeros((n, m) if m else n), dt)


# ---------------------------------------------------------------------------------------------------------------
# fields and SNodes
# ---------------------------------------------------------------------------------------------------------------
STRUCT_FOR_SORTED = True  # struct-for visits active cells in a fixed order (a legal schedule of the parallel loop; reproducible)
STRUCT_FOR_BLOCK = 16     # dense block edge used for the block-major order of 4-D fields


def _key(idx):
    if idx is None:
        return ()
    if isinstance(idx, Vec):
        return tuple(int(v) for v in idx.a)
    if isinstance(idx, (tuple, list, np.ndarray)):
        return tuple(int(v) for v in idx)
    return (int(idx),)


class Field:
    def __init__(self, dtype=None, shape=None, elem=()):
        self.dt = _npdt(dtype) or np.float32
        self.elem = tuple(elem)
        if shape is None:
            self.shape = None
        elif isinstance(shape, (tuple, list)):
            self.shape = tuple(int(s) for s in shape)
        else:
            self.shape = (int(shape),)
        self.d = {}
        self.snode = None

    def _is_float(self):
        return np.issubdtype(self.dt, np.floating)

    def _load(self, v):
        if self.elem:
            a = np.asarray(v)
            if a.dtype.kind == "f":
                return Vec(a.astype(np.float16 if self.dt == np.float16 else np.float32))
            return Vec(a.astype(np.int32))
        if self._is_float():
            return H16(v) if self.dt == np.float16 else np.float32(v)
        return int(v)

    def _store(self, v):
        if self.elem:
            a = _raw(v, np.float32 if self._is_float() else np.int64)
            return np.broadcast_to(np.asarray(a).astype(self.dt), self.elem).copy()
        if self._is_float():
            return self.dt(np.float32(float(v)))            # value converted to the field's type (f32 -> f16 rounds)
        return np.array(int(v)).astype(self.dt)[()]         # integer stores wrap like the C types do

    def _zero(self):
        return np.zeros(self.elem, self.dt) if self.elem else self.dt(0)

    def __getitem__(self, idx):
        k = _key(idx)
        v = self.d.get(k)
        if v is None:
            if self.elem and self.dt == np.float32:
                v = self.d[k] = self._zero()   # f32 vector cell: hand out the stored array (`f[idx][c] = x` writes through; activates)
            else:
                return self._load(self._zero())
        if self.elem and self.dt == np.float32:
            return Vec(v)                                   # write-through (self.input_R[None][i, j] = ...)
        return self._load(v)

    def __setitem__(self, idx, value):
        self.d[_key(idx)] = self._store(value)

    def __iter__(self):  # struct-for over the active cells
        ks = list(self.d.keys())
        if not STRUCT_FOR_SORTED:
            return iter(ks)
        if ks and len(ks[0]) == 4:  # (s, i, j, k): block-major like Taichi's pointer->dense traversal (16^3 blocks), then row-major
            B = STRUCT_FOR_BLOCK
            return iter(sorted(ks, key=lambda k: (k[0], k[1] // B, k[2] // B, k[3] // B, k[1] % B, k[2] % B, k[3] % B)))
        return iter(sorted(ks))

    def parent(self, n=1):
        if n not in (0, 1):  # 0: the field's own node, 1: its innermost container - one cell per element either way
            raise NotImplementedError("taichi_emu: struct-for over ancestor SNodes (level-of-detail export) is not emulated")
        return self

    def fill(self, v):
        for k in list(self.d.keys()):
            self.d[k] = self._store(v)

    def to_numpy(self):
        assert self.shape is not None, "to_numpy of a sparse field: read .d"
        out = np.zeros(self.shape + self.elem, self.dt)
        for k, v in self.d.items():
            out[k] = v
        return out

    def from_numpy(self, a):
        a = np.asarray(a)
        for k in np.ndindex(*self.shape):
            self.d[k] = self._store(a[k])


class SNode:
    def __init__(self, parent=None):
        self._parent, self.children, self.fields = parent, [], []

    def _child(self, *a, **k):
        c = SNode(self)
        self.children.append(c)
        return c

    pointer = dense = bitmasked = dynamic = hash = _child

    def place(self, *fields, offset=None, **k):
        for f in fields:
            if f is not None:
                self.fields.append(f)
                f.snode = self
        return self

    def parent(self, n=1):
        s = self
        for _ in range(n):
            s = s._parent if s._parent is not None else s
        return s

    def deactivate_all(self):
        for f in self.fields:
            f.d.clear()
        for c in self.children:
            c.deactivate_all()


# ---------------------------------------------------------------------------------------------------------------
# functions
# ---------------------------------------------------------------------------------------------------------------
def _identity_decorator(*a, **k):
    return a[0] if (len(a) == 1 and callable(a[0]) and not k) else (lambda f: f)


class _VecType:
    """ti.types.vector(n, dtype): a type token usable in @ti.dataclass annotations."""

    def __init__(self, n, dt):
        self.n, self.dt = n, dt

    def __call__(self, *a):
        return _vec(list(a[0]) if len(a) == 1 else list(a), self.dt)


def _zero_of(tp):
    if isinstance(tp, _VecType):
        return _vec(np.zeros(tp.n), tp.dt if tp.dt is not None else f32)
    npd = _npdt(tp)
    if npd is not None and np.issubdtype(npd, np.floating):
        return np.float32(0.0)
    return 0


class _StructField:
    """<Struct>.field(shape=n): elements are created on first access and handed out by reference."""

    def __init__(self, cls):
        self.cls, self.items = cls, {}

    def __getitem__(self, i):
        k = _key(i)
        it = self.items.get(k)
        if it is None:
            it = self.items[k] = self.cls()
        return it

    def __setitem__(self, i, v):
        self.items[_key(i)] = v


def ti_dataclass(cls):
    """@ti.dataclass: annotated members become zero-initialised attributes, @ti.func methods stay methods."""
    ann = dict(getattr(cls, "__annotations__", {}))

    def __init__(self, **kw):
        for name, tp in ann.items():
            setattr(self, name, kw.get(name, _zero_of(tp)))

    cls.__init__ = __init__
    cls.field = classmethod(lambda c, shape=None, **k: _StructField(c))
    return cls


def emu_atomic_add_(container, key, v):
    """ti.atomic_add(container[key], v) in expression position (rewritten by the loader)."""
    old = container[key]
    container[key] = old + v
    return old


def cast(v, dt):
    npd = _npdt(dt)
    if isinstance(v, Vec):
        a = v.a
        if np.issubdtype(npd, np.integer) and a.dtype.kind == "f":
            a = np.trunc(a.astype(np.float32))
        return Vec(a.astype(npd))
    x = float(v) if _cls(v) else int(v)
    if np.issubdtype(npd, np.integer):
        return int(np.trunc(x))
    return H16(x) if npd == np.float16 else np.float32(x)


def _round_half_away(a):
    """llvm.round: nearest integer, halves away from zero - exact (|a| + 0.5 would round 0.49999997 up in f32)."""
    a = np.asarray(a, dtype=np.float32)
    t = np.trunc(a)
    return np.where(np.abs(a - t) >= np.float32(0.5), t + np.copysign(np.float32(1.0), a), t).astype(np.float32)


def ti_round(v, dt=None):
    r = _round_half_away(v.a if isinstance(v, Vec) else float(v))
    out = Vec(r) if r.ndim else np.float32(r)
    return cast(out, dt) if dt is not None else out


def ti_floor(v, dt=None):
    r = np.floor(np.asarray(v.a if isinstance(v, Vec) else float(v), dtype=np.float32))
    out = Vec(r) if r.ndim else np.float32(r)
    return cast(out, dt) if dt is not None else out


def ti_abs(v):
    if isinstance(v, Vec):
        return Vec(np.abs(v.a))
    if isinstance(v, H16):
        return abs(v)
    return np.float32(abs(v)) if _cls(v) else abs(int(v))


def ti_sqrt(v):
    if isinstance(v, H16):
        return H16(np.sqrt(v.v))
    return np.float32(np.sqrt(np.float32(float(v))))


def _strong(x):
    return np.float32(x) if isinstance(x, float) else x


def ti_min(a, b):
    a, b = _strong(a), _strong(b)
    r = a if a < b else b
    c = max(_cls(a), _cls(b))
    return cast(r, {0: i32, 1: f16, 2: f32}[c]) if c else int(r)


def ti_max(a, b):
    a, b = _strong(a), _strong(b)
    r = a if a > b else b
    c = max(_cls(a), _cls(b))
    return cast(r, {0: i32, 1: f16, 2: f32}[c]) if c else int(r)


def ti_static(*a):
    return a[0] if len(a) == 1 else a


class I32(int):
    """A RUNTIME i32 value (loop index).  int (+) Python float happens in f32 in a Taichi kernel (the float is an f32
    constant there), not in Python's f64: `_len = _j * self.voxel_scale` (mapping_common.py:173) is an f32 product."""

    def _f(self, o, op):
        if isinstance(o, float) and not isinstance(o, np.floating):
            return op(np.float32(int(self)), np.float32(o))
        return NotImplemented

    def __mul__(self, o):
        r = self._f(o, np.multiply)
        return r if r is not NotImplemented else (I32(int(self) * int(o)) if isinstance(o, int) and not isinstance(o, bool) else int.__mul__(self, o) if isinstance(o, int) else NotImplemented)

    __rmul__ = __mul__

    def __add__(self, o):
        r = self._f(o, np.add)
        return r if r is not NotImplemented else (I32(int(self) + int(o)) if isinstance(o, int) else NotImplemented)

    __radd__ = __add__

    def __sub__(self, o):
        r = self._f(o, np.subtract)
        return r if r is not NotImplemented else (I32(int(self) - int(o)) if isinstance(o, int) else NotImplemented)

    def __rsub__(self, o):
        if isinstance(o, float) and not isinstance(o, np.floating):
            return np.float32(o) - np.float32(int(self))
        return I32(int(o) - int(self)) if isinstance(o, int) else NotImplemented

    def __truediv__(self, o):
        if isinstance(o, (int, float)) and not isinstance(o, np.floating):
            return np.float32(int(self)) / np.float32(o)
        return NotImplemented

    def __rtruediv__(self, o):
        if isinstance(o, (int, float)) and not isinstance(o, np.floating):
            return np.float32(o) / np.float32(int(self))
        return NotImplemented


def ti_range(*args):
    """range() with Taichi's implicit float -> int conversion of the bounds (truncation); yields runtime i32 values."""
    return (I32(v) for v in range(*[int(a) for a in args]))


def ti_sign(val):  # mapping_common.py:5-7 `(0 < val) - (val < 0)`
    return int(0 < val) - int(val < 0)


def _atomic_add_unreachable(*a, **k):
    raise RuntimeError("ti.atomic_add must have been rewritten by the loader")


def make_ti():
    ti = types.ModuleType("taichi")
    ti.__dict__.update(dict(
        f16=f16, f32=f32, f64=f64, i8=i8, i16=i16, i32=i32, i64=i64, u8=u8, u16=u16, u32=u32, int32=int32, float32=float32,
        kernel=_identity_decorator, func=_identity_decorator, data_oriented=_identity_decorator, dataclass=ti_dataclass,
        static=ti_static, template=lambda *a, **k: None,
        types=types.SimpleNamespace(ndarray=lambda *a, **k: None, vector=lambda n, dt=None: _VecType(n, dt), matrix=lambda *a, **k: None),
        Vector=Vector, Matrix=Matrix, field=lambda dtype=None, shape=None, **k: Field(dtype, shape=shape),
        root=SNode(), i="i", j="j", k="k", l="l", ij="ij", ijk="ijk", ijkl="ijkl",
        cast=cast, abs=ti_abs, sqrt=ti_sqrt, round=ti_round, floor=ti_floor,
        min=ti_min, max=ti_max, atomic_add=_atomic_add_unreachable, grouped=lambda it: (_vec(list(kk), i32) for kk in it),
        loop_config=lambda **k: None, init=lambda **k: None, cuda="cuda", cpu="cpu", gpu="gpu",
        random=lambda dtype=float: 0,
    ))
    return ti


# ---------------------------------------------------------------------------------------------------------------
# loader: import the reference's files with `ti.atomic_add` rewritten and `range` shadowed
# ---------------------------------------------------------------------------------------------------------------
class _AtomicRewriter(ast.NodeTransformer):
    @staticmethod
    def _is_atomic(call):
        return (isinstance(call, ast.Call) and isinstance(call.func, ast.Attribute) and call.func.attr == "atomic_add" and
                isinstance(call.func.value, ast.Name) and call.func.value.id == "ti")

    def visit_Call(self, node):
        self.generic_visit(node)
        if self._is_atomic(node) and isinstance(node.args[0], ast.Subscript):  # works in any expression position
            tgt = node.args[0]
            return ast.Call(func=ast.Name(id="emu_atomic_add_", ctx=ast.Load()), args=[tgt.value, tgt.slice, node.args[1]], keywords=[])
        return node

    def visit_Assign(self, node):
        self.generic_visit(node)
        if self._is_atomic(node.value) and len(node.targets) == 1:
            x, v = node.value.args
            x_store = ast.parse(ast.unparse(x)).body[0].value
            self._set_ctx(x_store, ast.Store())
            return [ast.Assign(targets=node.targets, value=x),
                    ast.Assign(targets=[x_store], value=ast.BinOp(left=ast.parse(ast.unparse(x)).body[0].value, op=ast.Add(), right=v))]
        return node

    def visit_Expr(self, node):
        self.generic_visit(node)
        if self._is_atomic(node.value):
            x, v = node.value.args
            x_store = ast.parse(ast.unparse(x)).body[0].value
            self._set_ctx(x_store, ast.Store())
            return ast.Assign(targets=[x_store], value=ast.BinOp(left=x, op=ast.Add(), right=v))
        return node

    @staticmethod
    def _set_ctx(node, ctx):
        if isinstance(node, (ast.Name, ast.Subscript, ast.Attribute)):
            node.ctx = ctx


class _RefLoader(importlib.machinery.SourceFileLoader):
    def source_to_code(self, data, path, *, _optimize=-1):
        tree = ast.parse(data, filename=path)
        tree = _AtomicRewriter().visit(tree)
        tree.body.insert(0, ast.parse("from oracle.taichi_emu import ti_range as range, emu_atomic_add_").body[0])
        ast.fix_missing_locations(tree)
        return compile(tree, path, "exec", dont_inherit=True, optimize=_optimize)


class _RefFinder(importlib.abc.MetaPathFinder):
    def __init__(self, root):
        self.root = root

    def find_spec(self, fullname, path=None, target=None):
        if fullname != "taichi_slam" and not fullname.startswith("taichi_slam."):
            return None
        rel = fullname.split(".")
        base = os.path.join(self.root, *rel)
        if os.path.isdir(base):
            init = os.path.join(base, "__init__.py")
            return importlib.util.spec_from_file_location(fullname, init, loader=_RefLoader(fullname, init), submodule_search_locations=[base])
        if os.path.exists(base + ".py"):
            return importlib.util.spec_from_file_location(fullname, base + ".py", loader=_RefLoader(fullname, base + ".py"))
        return None


def load_reference(root="/root/reference"):
    """Import <root>/taichi_slam/mapping/{mapping_common,dense_tsdf,taichi_octomap,marching_cube_mesher}.py through the
    stand-in.  Must run in a process that has not imported this repository's own `taichi_slam` alias package."""
    from unittest.mock import MagicMock
    assert "taichi_slam" not in sys.modules, "load the reference in a fresh process"
    sys.modules["taichi"] = make_ti()
    for m in ("matplotlib", "matplotlib.cm", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.mplot3d", "lcm", "transformations"):
        if m not in sys.modules:
            try:
                __import__(m)
            except Exception:
                sys.modules[m] = MagicMock()
    if isinstance(sys.modules["matplotlib"], MagicMock):  # init_colormap (mapping_common.py:158-163) stores cm.jet(x)[0:3]
        cm = types.ModuleType("matplotlib.cm")
        cm.jet = lambda x: (float(min(max(1.5 - abs(4 * x - 3), 0), 1)), float(min(max(1.5 - abs(4 * x - 2), 0), 1)),
                            float(min(max(1.5 - abs(4 * x - 1), 0), 1)), 1.0)
        sys.modules["matplotlib.cm"] = cm
        sys.modules["matplotlib"].cm = cm
    sys.meta_path.insert(0, _RefFinder(root))
    # mapping/__init__.py star-imports every module incl. topo_graph / submap_mapping; import the kernels' modules directly
    pkg = types.ModuleType("taichi_slam")
    pkg.__path__ = [os.path.join(root, "taichi_slam")]
    sys.modules["taichi_slam"] = pkg
    sub = types.ModuleType("taichi_slam.mapping")
    sub.__path__ = [os.path.join(root, "taichi_slam", "mapping")]
    sys.modules["taichi_slam.mapping"] = sub
    mods = {}
    for name in ("mapping_common", "dense_tsdf", "taichi_octomap", "marching_cube_mesher", "topo_graph"):
        mod = importlib.import_module("taichi_slam.mapping." + name)
        if hasattr(mod, "sign"):
            mod.sign = ti_sign
        mods[name] = mod
    return types.SimpleNamespace(**mods)
