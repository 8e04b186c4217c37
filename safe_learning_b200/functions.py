"""Function objects of the region-of-attraction path, as parameter containers for libslb200.

API surface follows ``safe_learning/functions.py`` of the reference (class names, constructor
arguments, attributes); cited line numbers are relative to /root/reference.  Where the
reference's objects emit TF1 graph nodes, these objects (a) describe themselves to the CUDA
kernels through an ``slb_function`` / ``slb_gp_stack`` descriptor and (b) evaluate eagerly on
the GPU when called with numpy arrays, returning numpy arrays.  There is no CPU path.

gpflow is replaced by small gpflow-free containers: the ``kernels`` (RBF, Matern12/32/52, Linear,
Constant, White, sums and products, ``active_dims``; arithmetic of ``gpflow==0.4.0``
``kernels``) and ``GPRCached`` (``functions.py:357-458``).
"""

from __future__ import annotations

import hashlib
import itertools

import numpy as np
import scipy.signal
import scipy.spatial
import torch

from . import _device as dev
from . import _native as nat
from .configuration import Configuration

config = Configuration()

__all__ = ["DimensionError", "GridWorld", "Function", "DeterministicFunction",
           "UncertainFunction", "ConstantFunction", "LinearSystem", "QuadraticFunction",
           "Saturation", "AbsFunction", "Norm1Function", "MaxAbsFunction", "ScaledFunction",
           "Triangulation", "TriangulationGradient",
           "Kernel", "RBF", "Matern12", "Matern32", "Matern52", "Linear", "Constant", "Bias",
           "White", "Sum", "Add", "Product", "Prod", "kernels", "Likelihood", "GPRCached", "GPR",
           "GaussianProcess", "FunctionStack",
           "InvertedPendulum", "CartPole", "LyapunovNetwork", "NeuralNetwork",
           "concatenate_inputs"]


class DimensionError(Exception):
    """``functions.py:575-576``."""


def concatenate_inputs(inputs):
    """Column-concatenate call arguments ([x, u] layout; ``utilities.py:123-159``)."""
    cols = [np.atleast_2d(np.asarray(a, dtype=np.float64)) for a in inputs]
    return cols[0] if len(cols) == 1 else np.hstack(cols)


# =============================================================================== GridWorld
class GridWorld(object):
    """Regular grid (``functions.py:579-817``).

    The kernels never read coordinates from memory: ``descriptor()`` carries
    offset / unit_maxes / num_points and every thread rebuilds ``ijk * unit_maxes + offset``
    (``:731``) from its flat index.  The host-side helpers below are vectorised numpy index
    arithmetic (not a compute path).
    """

    def __init__(self, limits, num_points):
        self.limits = np.atleast_2d(limits).astype(np.float64)
        npts = np.broadcast_to(num_points, len(self.limits))
        self.num_points = npts.astype(np.int64, copy=True)
        if np.any(self.num_points < 2):
            raise DimensionError("There must be at least 2 points in each dimension.")
        if len(self.limits) > nat.SLB_MAX_DIM:
            raise DimensionError("at most %d grid dimensions are supported" % nat.SLB_MAX_DIM)
        self.offset = self.limits[:, 0]
        self.unit_maxes = (self.limits[:, 1] - self.offset) / (self.num_points - 1)
        self.offset_limits = np.stack((np.zeros(len(self.limits)),
                                       self.limits[:, 1] - self.offset), axis=1)
        self.discrete_points = [np.linspace(lo, hi, n, dtype=np.float64)
                                for (lo, hi), n in zip(self.limits, self.num_points)]
        self.nrectangles = int(np.prod(self.num_points - 1))
        self.nindex = int(np.prod(self.num_points))
        self.ndim = len(self.limits)
        self._all_points = None
        self._points_dev = None

    def __len__(self):
        return self.nindex

    # ---- descriptor for the kernels
    def descriptor(self, need_points=False):
        g = nat.SlbGrid()
        g.ndim = self.ndim
        g.nindex = self.nindex
        for c in range(self.ndim):
            g.num_points[c] = int(self.num_points[c])
            g.offset[c] = float(self.offset[c])
            g.unit_maxes[c] = float(self.unit_maxes[c])
            g.upper[c] = float(self.limits[c, 1])
        if need_points:
            if self._points_dev is None:
                self._points_dev = dev.to_device(np.concatenate(self.discrete_points))
            g.discrete_points = self._points_dev.data_ptr()
        return g

    # ---- reference API
    @property
    def all_points(self):
        """All grid points, C order, last dimension fastest (``:622-638``).  Host array; the
        sweeps do not use it."""
        if self._all_points is None:
            mesh = np.meshgrid(*self.discrete_points, indexing="ij")
            self._all_points = np.column_stack([m.ravel() for m in mesh])
        return self._all_points

    def sample_continuous(self, num_samples):
        rand = np.random.uniform(0, 1, size=(num_samples, self.ndim))
        return rand * np.diff(self.limits, axis=1).T + self.offset

    def sample_discrete(self, num_samples, replace=False):
        idx = np.random.choice(self.nindex, size=num_samples, replace=replace)
        return self.index_to_state(idx)

    def _check_dimensions(self, states):
        if not states.shape[1] == self.ndim:
            raise DimensionError("the input argument has the wrong dimensions.")

    def _center_states(self, states, clip=True):
        eps = np.finfo(np.float64).eps
        states = np.atleast_2d(states).astype(np.float64) - self.offset[None, :]
        if clip:
            np.clip(states, self.offset_limits[:, 0] + 2 * eps,
                    self.offset_limits[:, 1] - 2 * eps, out=states)
        return states

    def index_to_state(self, indices):
        ijk = np.stack(np.unravel_index(np.atleast_1d(indices), self.num_points), axis=1)
        return ijk.astype(np.float64) * self.unit_maxes + self.offset

    def state_to_index(self, states):
        states = np.atleast_2d(states)
        self._check_dimensions(states)
        clipped = np.clip(states, self.limits[:, 0], self.limits[:, 1])
        ijk = np.rint((clipped - self.offset) * (1. / self.unit_maxes)).astype(np.int32)
        return np.ravel_multi_index(ijk.T, self.num_points)

    def state_to_rectangle(self, states):
        cells = []
        for i, (pts, n) in enumerate(zip(self.discrete_points, self.num_points)):
            cells.append(np.clip(np.digitize(states[:, i], pts) - 1, 0, n - 2))
        return np.ravel_multi_index(cells, self.num_points - 1)

    def rectangle_to_state(self, rectangles):
        ijk = np.stack(np.unravel_index(np.atleast_1d(rectangles), self.num_points - 1), axis=1)
        return ijk.astype(np.float64) * self.unit_maxes + self.offset

    def rectangle_corner_index(self, rectangles):
        ijk = np.vstack(np.unravel_index(rectangles, self.num_points - 1))
        return np.ravel_multi_index(np.atleast_2d(ijk), self.num_points)


# =============================================================================== Function base
class Function(object):
    """Base class (``functions.py:31-122``): ``fun(*inputs)`` concatenates the inputs and
    evaluates on the GPU; ``+``/``*``/``-`` with scalars build fused wrappers."""

    input_dim = None
    output_dim = None

    def __init__(self, name="function"):
        self.name = name
        self.feed_dict = {}

    # descriptor protocol --------------------------------------------------------------
    def descriptor(self):
        """Return an ``slb_function``; device buffers it points to are owned by ``self``."""
        raise NotImplementedError("%s cannot be fused into the CUDA kernels"
                                  % type(self).__name__)

    @property
    def parameters(self):
        return []

    @property
    def version(self):
        """Changes whenever the descriptor would change (callers cache descriptors on it).
        Plain parameter containers are immutable after construction."""
        return 0

    # eager evaluation -----------------------------------------------------------------
    def __call__(self, *inputs):
        return self.evaluate_device(concatenate_inputs(inputs)).cpu().numpy()

    def evaluate_device(self, points):
        """points: numpy [n, in] or device tensor -> device tensor [n, out]."""
        lib = nat.load()
        desc = self.descriptor()
        pts = dev.to_device(points)
        if pts.dim() != 2 or pts.shape[1] != desc.in_dim:
            raise DimensionError("%s expects %d input columns, got shape %s"
                                 % (type(self).__name__, desc.in_dim, tuple(pts.shape)))
        ncols = 1 if (desc.flags & (nat.FLAG_NORM1 | nat.FLAG_MAXABS)
                      or desc.kind == nat.FN_QUADRATIC) else desc.out_dim
        out = dev.empty((pts.shape[0], ncols))
        nat.check(lib.slb_eval_function(dev.stream(), desc, pts.data_ptr(), pts.shape[0],
                                        out.data_ptr()), "slb_eval_function")
        return out

    # differentiable application (torch autograd) --------------------------------------------
    def jacobian_device(self, points):
        """d out / d in at device points [n, in] -> device tensor [n, out, in] (the reference gets
        these from ``tf.gradients``; here every fusable object states its own)."""
        raise NotImplementedError("%s has no device Jacobian" % type(self).__name__)

    def torch(self, points):
        """``fun(points)`` on a device tensor [n, in] as a node of torch's autograd graph: forward
        = the fused CUDA evaluation, backward = ``grad_out @ jacobian_device(points)`` on the
        device (``PolicyIteration.future_values`` with tensors, ``reinforcement_learning.py:65-114``
        under ``tf.gradients`` in ``examples/inverted_pendulum.ipynb`` cell 17)."""
        return _FusedApply.apply(points, self)

    # algebra (``functions.py:112-122``) --------------------------------------------------
    def __neg__(self):
        return ScaledFunction(self, -1.0)

    def __mul__(self, other):
        if np.isscalar(other):
            return ScaledFunction(self, float(other))
        raise NotImplementedError("only multiplication by a scalar is fused on the GPU")

    __rmul__ = __mul__

    def __abs__(self):
        return AbsFunction(self)

    def __add__(self, other):
        raise NotImplementedError("AddedFunction (functions.py:125-160) is outside the fused "
                                  "hot path of this build")


class _FusedApply(torch.autograd.Function):
    """Fused CUDA evaluation of a Function object inside torch's autograd graph."""

    @staticmethod
    def forward(ctx, points, fun):
        points = points.detach().contiguous()
        ctx.fun = fun
        ctx.save_for_backward(points)
        return fun.evaluate_device(points)

    @staticmethod
    def backward(ctx, grad_out):
        (points,) = ctx.saved_tensors
        jac = ctx.fun.jacobian_device(points)                       # [n, out, in]
        return torch.einsum("no,noi->ni", grad_out.contiguous(), jac), None


class DeterministicFunction(Function):
    """``functions.py:233-238``."""


class UncertainFunction(Function):
    """``functions.py:202-230``."""


class ConstantFunction(DeterministicFunction):
    """``functions.py:241-251``."""

    def __init__(self, constant, input_dim=1, name="constant_function"):
        super().__init__(name)
        self.constant = np.atleast_1d(np.asarray(constant, dtype=np.float64)).ravel()
        self.input_dim, self.output_dim = input_dim, len(self.constant)

    def descriptor(self):
        d = nat.SlbFunction()
        d.kind, d.in_dim, d.out_dim = nat.FN_CONSTANT, self.input_dim, self.output_dim
        for i, v in enumerate(self.constant):
            d.cparams[i] = float(v)
        return d

    def jacobian_device(self, points):
        return dev.zeros((points.shape[0], self.output_dim, self.input_dim))


class LinearSystem(DeterministicFunction):
    """``y = [x, u] A^T`` (``functions.py:1546-1583``)."""

    def __init__(self, matrices, name="linear_system"):
        super().__init__(name)
        if isinstance(matrices, np.ndarray):
            matrices = (matrices,)
        self.matrix = np.hstack([np.atleast_2d(m).astype(np.float64) for m in matrices])
        self.output_dim, self.input_dim = self.matrix.shape
        self._matrix_dev = None

    def descriptor(self):
        if self._matrix_dev is None:
            self._matrix_dev = dev.to_device(self.matrix)
        d = nat.SlbFunction()
        d.kind, d.in_dim, d.out_dim = nat.FN_LINEAR, self.input_dim, self.output_dim
        d.matrix = self._matrix_dev.data_ptr()
        return d

    def jacobian_device(self, points):
        self.descriptor()
        return self._matrix_dev.unsqueeze(0).expand(points.shape[0], -1, -1)


class QuadraticFunction(DeterministicFunction):
    """``sum((x P) * x)`` with P as given, not symmetrised (``functions.py:1513-1543``)."""

    def __init__(self, matrix, name="quadratic"):
        super().__init__(name)
        self.matrix = np.atleast_2d(matrix).astype(np.float64)
        self.ndim = self.matrix.shape[0]
        self.input_dim, self.output_dim = self.ndim, 1
        self._matrix_dev = None

    def descriptor(self):
        if self._matrix_dev is None:
            self._matrix_dev = dev.to_device(self.matrix)
        d = nat.SlbFunction()
        d.kind, d.in_dim, d.out_dim = nat.FN_QUADRATIC, self.ndim, 1
        d.matrix = self._matrix_dev.data_ptr()
        return d

    def gradient(self, points=None):
        """``x (P + P^T)`` as a fusable LinearSystem (``:1541-1543``); evaluated when points
        are given."""
        grad = LinearSystem((self.matrix + self.matrix.T).T)
        return grad if points is None else grad(points)

    def jacobian_device(self, points):
        self.descriptor()
        sym = self._matrix_dev + self._matrix_dev.T
        return (points @ sym).unsqueeze(1)                          # x (P + P^T)


class _PostOp(DeterministicFunction):
    """A post-operation fused onto a wrapped function's descriptor.  The kernels apply
    saturate -> abs -> norm1 -> scale in that order (``slb200.h``), so wrappers must be
    nested in that order (``MaxAbsFunction`` takes the place of norm1)."""

    _order = 0

    def __init__(self, fun, name):
        super().__init__(name)
        if not isinstance(fun, Function):
            raise TypeError("%s wraps a Function object, got %r" % (type(self).__name__, fun))
        inner = getattr(fun, "_order", 0)
        if inner >= self._order:
            raise NotImplementedError(
                "%s around %s: post-operations fuse only in the order "
                "saturate -> abs -> norm1 -> scale" % (type(self).__name__, type(fun).__name__))
        self.fun = fun
        self.input_dim = fun.input_dim

    @property
    def parameters(self):
        return self.fun.parameters

    @property
    def version(self):
        return (id(self.fun), self.fun.version, getattr(self, "lower", None),
                getattr(self, "upper", None), getattr(self, "factor", None))


class Saturation(_PostOp):
    """``min(max(fun(x), lower), upper)`` (``functions.py:310-354``)."""

    _order = 1

    def __init__(self, fun, lower, upper, name="saturation"):
        super().__init__(fun, name)
        self.lower, self.upper = float(lower), float(upper)
        self.output_dim = fun.output_dim

    def descriptor(self):
        d = self.fun.descriptor()
        d.flags |= nat.FLAG_SATURATE
        d.lower, d.upper = self.lower, self.upper
        return d

    def jacobian_device(self, points):
        inner = self.fun.evaluate_device(points)
        free = ((inner > self.lower) & (inner < self.upper)).to(torch.float64)
        return self.fun.jacobian_device(points) * free.unsqueeze(2)   # tf.clip_by_value's gradient


class AbsFunction(_PostOp):
    """``|fun(x)|`` element-wise -- the per-dimension Lipschitz lambda
    ``tf.abs(grad_lyapunov_function(x))`` of ``examples/adaptive_safety_verification.ipynb``
    cell 17, as a fusable object."""

    _order = 2

    def __init__(self, fun, name="abs"):
        super().__init__(fun, name)
        self.output_dim = fun.output_dim

    def descriptor(self):
        d = self.fun.descriptor()
        d.flags |= nat.FLAG_ABS
        return d

    def jacobian_device(self, points):
        sign = torch.sign(self.fun.evaluate_device(points))
        return self.fun.jacobian_device(points) * sign.unsqueeze(2)


class Norm1Function(_PostOp):
    """``tf.norm(fun(x), ord=1, axis=1, keepdims=True)`` (same notebook cell)."""

    _order = 3

    def __init__(self, fun, name="norm1"):
        super().__init__(fun, name)
        self.output_dim = 1

    def descriptor(self):
        d = self.fun.descriptor()
        d.flags |= nat.FLAG_NORM1
        return d

    def jacobian_device(self, points):
        sign = torch.sign(self.fun.evaluate_device(points))
        return (self.fun.jacobian_device(points) * sign.unsqueeze(2)).sum(dim=1, keepdim=True)


class MaxAbsFunction(_PostOp):
    """``tf.reduce_max(tf.abs(fun(x)), axis=1, keepdims=True)`` -- the Lipschitz lambda of
    ``examples/inverted_pendulum.ipynb`` cell 14 (around ``value_function.gradient``), as a
    fusable object."""

    _order = 3

    def __init__(self, fun, name="maxabs"):
        super().__init__(fun, name)
        self.output_dim = 1

    def descriptor(self):
        d = self.fun.descriptor()
        d.flags |= nat.FLAG_MAXABS
        return d


class ScaledFunction(_PostOp):
    """``fun * c`` / ``-fun`` (``MultipliedFunction`` with a constant, ``functions.py:163-199``)."""

    _order = 4

    def __init__(self, fun, factor, name="scaled"):
        super().__init__(fun, name)
        self.factor = float(factor)
        self.output_dim = fun.output_dim

    def __getattr__(self, item):          # e.g. (-value_function).discretization
        if item in ("fun", "factor"):
            raise AttributeError(item)
        return getattr(self.fun, item)

    def descriptor(self):
        d = self.fun.descriptor()
        d.flags |= nat.FLAG_SCALE
        d.out_scale = self.factor
        return d

    def jacobian_device(self, points):
        return self.fun.jacobian_device(points) * self.factor


# =============================================================================== Triangulation
class _Delaunay1D(object):
    """Two-point stand-in for scipy's Delaunay in 1-D (``functions.py:935-978``)."""

    def __init__(self, points):
        self.points = points
        self.nsimplex = 1
        self.simplices = np.array([[0, 1]])


class _TriangulationTables(object):
    """Host-side tables of ``_Triangulation`` (``functions.py:1002-1101``): the unit
    hyper-rectangle is triangulated once by Qhull; the kernels take the resulting
    ``unit_simplices`` / ``hyperplanes`` instead of assuming a particular split."""

    def __init__(self, discretization, project=False):
        self.discretization = disc = discretization
        self.input_dim = disc.ndim
        self.project = project
        if disc.ndim == 1:
            self.triangulation = _Delaunay1D(np.array([[0.0], [disc.unit_maxes[0]]]))
        else:
            corners = np.array(list(itertools.product(*np.diag(disc.unit_maxes))))
            self.triangulation = scipy.spatial.Delaunay(corners)
        mapping = disc.state_to_index(np.atleast_2d(self.triangulation.points) + disc.offset)
        self.unit_simplices = mapping[np.asarray(self.triangulation.simplices)].astype(np.int64)
        self.nsimplex_unit = int(self.triangulation.nsimplex)
        self.nsimplex = self.nsimplex_unit * disc.nrectangles
        self.hyperplanes = np.empty((self.nsimplex_unit, disc.ndim, disc.ndim))
        for i, simplex in enumerate(self.unit_simplices):
            pts = disc.index_to_state(simplex)
            self.hyperplanes[i] = np.linalg.inv(pts[1:] - pts[:1])
        # Queries clipped in every dimension land on a unit-cell corner shared by several
        # simplices; which one Qhull's walk returns decides the (discontinuous) extrapolation
        # of a non-projected query, so ask Qhull once per corner pattern (functions.py:1120-1124).
        self.corner_simplex = np.zeros(2 ** disc.ndim, dtype=np.int32)
        if disc.ndim > 1:
            eps = np.finfo(np.float64).eps
            lo = disc.offset_limits[:, 0] + 2 * eps
            hi = disc.offset_limits[:, 1] - 2 * eps
            for pattern in range(2 ** disc.ndim):
                bits = np.array([(pattern >> c) & 1 for c in range(disc.ndim)], dtype=bool)
                unit = np.where(bits, hi, lo) % disc.unit_maxes
                self.corner_simplex[pattern] = int(self.triangulation.find_simplex(unit[None, :])[0])

    @property
    def limits(self):
        return self.discretization.limits

    @property
    def nindex(self):
        return self.discretization.nindex


class Triangulation(DeterministicFunction):
    """Piecewise-linear interpolation on a GridWorld (``functions.py:1372-1510``).

    The vertex values live in HBM (``_param_dev`` [nindex, out]); ``parameters`` exposes
    them like the reference's single tf.Variable: ``tri.parameters[0]`` is the [nindex, out]
    array, and assigning ``tri.parameters = values`` re-uploads.
    """

    def __init__(self, discretization, vertex_values, project=False, name="triangulation"):
        super().__init__(name)
        self.tri = _TriangulationTables(discretization, project=project)
        self.input_dim = self.tri.input_dim
        self._param_dev = None
        self._hyper_dev = None
        self._simp_dev = None
        self._corner_dev = None
        self._version = 0
        self.output_dim = None
        if vertex_values is not None:
            self.parameters = vertex_values

    @property
    def project(self):
        return self.tri.project

    @project.setter
    def project(self, value):
        self.tri.project = value

    @property
    def discretization(self):
        return self.tri.discretization

    @property
    def nindex(self):
        return self.tri.nindex

    @property
    def parameters(self):
        if self._param_dev is None:
            return []
        return [self._param_dev.cpu().numpy()]

    @parameters.setter
    def parameters(self, values):
        if isinstance(values, (list, tuple)) and len(values) == 1:
            values = values[0]
        if isinstance(values, torch.Tensor):
            vals = values.to(dtype=torch.float64).reshape(self.nindex, -1)
            self._param_dev = vals.to(dev.device()).contiguous().clone()
        else:
            vals = np.asarray(values, dtype=np.float64).reshape(self.nindex, -1)
            self._param_dev = dev.to_device(vals)
        self.output_dim = int(self._param_dev.shape[1])
        self._version += 1

    @property
    def version(self):
        # the vertex table is swapped (new device buffer) by value_iteration, so its address
        # is part of the descriptor identity
        return (self._version, self.project,
                0 if self._param_dev is None else self._param_dev.data_ptr())

    def descriptor(self):
        if self._param_dev is None:
            raise ValueError("Triangulation has no vertex values")
        if self.input_dim > nat.SLB_MAX_DIM:
            raise DimensionError("Triangulation supports up to %d dims" % nat.SLB_MAX_DIM)
        if self._hyper_dev is None:
            self._hyper_dev = dev.to_device(self.tri.hyperplanes)
            self._simp_dev = dev.to_device(self.tri.unit_simplices, torch.int64)
            self._corner_dev = dev.to_device(self.tri.corner_simplex, torch.int32)
        d = nat.SlbFunction()
        d.kind, d.in_dim, d.out_dim = nat.FN_TRIANGULATION, self.input_dim, self.output_dim
        d.flags = nat.FLAG_PROJECT if self.project else 0
        d.matrix = self._param_dev.data_ptr()
        d.hyperplanes = self._hyper_dev.data_ptr()
        d.unit_simplices = self._simp_dev.data_ptr()
        d.corner_simplex = self._corner_dev.data_ptr() if self.input_dim > 1 else None
        d.nsimplex = self.tri.nsimplex_unit
        d.grid = self.discretization.descriptor(need_points=True)
        return d

    def gradient_function(self):
        """The gradient of the interpolant as a fusable function object (one value column)."""
        return TriangulationGradient(self)

    def gradient(self, points):
        """``Triangulation.gradient`` (``functions.py:1302-1326, 1506-1510``): the partial
        derivatives of the piecewise-linear interpolant, numpy ``[n, d]``."""
        return self.gradient_function()(points)

    def jacobian_device(self, points):
        if self.output_dim != 1:
            raise NotImplementedError("Jacobian of a multi-column Triangulation")
        grad = self.gradient_function().evaluate_device(points)       # [n, d]
        if self.project:      # clipped coordinates carry no gradient (tf.clip_by_value, :1479-1485)
            lim = dev.to_device(np.asarray(self.discretization.limits, dtype=np.float64))
            grad = grad * ((points >= lim[:, 0]) & (points <= lim[:, 1])).to(torch.float64)
        return grad.unsqueeze(1)


class TriangulationGradient(DeterministicFunction):
    """``x -> d Triangulation(x) / dx`` (piecewise constant; ``functions.py:1260-1326``), evaluated
    by the same simplex lookup as the value (``SLB_FLAG_GRADIENT``)."""

    def __init__(self, triangulation, name="triangulation_gradient"):
        super().__init__(name)
        if not isinstance(triangulation, Triangulation):
            raise TypeError("TriangulationGradient wraps a Triangulation")
        self.triangulation = triangulation
        self.input_dim = self.output_dim = triangulation.input_dim

    @property
    def parameters(self):
        return self.triangulation.parameters

    @property
    def version(self):
        return ("grad", self.triangulation.version)

    def descriptor(self):
        d = self.triangulation.descriptor()
        if d.out_dim != 1:
            raise DimensionError("the fused gradient needs a Triangulation with one value column")
        d.flags |= nat.FLAG_GRADIENT
        d.out_dim = self.input_dim
        return d


# =============================================================================== plants
def _pack_norm(cp, normalization, ns, base):
    """cparams[base:] = Tx[ns], Tu, 1/Tx[ns]; returns the flag value."""
    if normalization is None:
        return 0.0
    tx, tu = (np.asarray(n, dtype=np.float64).ravel() for n in normalization)
    inv = tx ** -1
    for i in range(ns):
        cp[base + i] = float(tx[i])
    cp[base + ns] = float(tu[0])
    for i in range(ns):
        cp[base + ns + 1 + i] = float(inv[i])
    return 1.0


class InvertedPendulum(DeterministicFunction):
    """Normalised inverted pendulum, 10 explicit-Euler sub-steps
    (``examples/utilities.py:144-289``)."""

    def __init__(self, mass, length, friction=0.0, dt=1 / 80, normalization=None,
                 name="inverted_pendulum"):
        super().__init__(name)
        self.mass, self.length, self.friction, self.dt = mass, length, friction, dt
        self.gravity = 9.81
        self.normalization = normalization
        if normalization is not None:
            self.normalization = [np.array(n, dtype=np.float64) for n in normalization]
            self.inv_norm = [n ** -1 for n in self.normalization]
        self.input_dim, self.output_dim = 3, 2

    @property
    def inertia(self):
        return self.mass * self.length ** 2

    def linearize(self):
        """Discretised, normalised linearisation (``examples/utilities.py:207-240``)."""
        g, l, b, inertia = self.gravity, self.length, self.friction, self.inertia
        A = np.array([[0, 1], [g / l, -b / inertia]], dtype=np.float64)
        B = np.array([[0], [1 / inertia]], dtype=np.float64)
        if self.normalization is not None:
            Tx, Tu = map(np.diag, self.normalization)
            Tx_inv, Tu_inv = map(np.diag, self.inv_norm)
            A = np.linalg.multi_dot((Tx_inv, A, Tx))
            B = np.linalg.multi_dot((Tx_inv, B, Tu))
        sysd = scipy.signal.StateSpace(A, B, np.eye(2), np.zeros((2, 1))).to_discrete(self.dt)
        return sysd.A, sysd.B

    def descriptor(self):
        d = nat.SlbFunction()
        d.kind, d.in_dim, d.out_dim = nat.FN_PENDULUM, 3, 2
        cp = d.cparams
        cp[0] = self.gravity / self.length
        cp[1] = self.inertia
        cp[2] = self.friction / self.inertia
        cp[3] = self.dt / 10
        cp[9] = _pack_norm(cp, self.normalization, 2, 4)     # [4..5] Tx, [6] Tu, [7..8] 1/Tx
        cp[10] = 1.0 if self.friction > 0 else 0.0
        return d


class CartPole(DeterministicFunction):
    """Cart-pole (``examples/utilities.py:292-437``)."""

    def __init__(self, pendulum_mass, cart_mass, length, rot_friction=0.0, dt=0.01,
                 normalization=None, name="CartPole"):
        super().__init__(name)
        self.pendulum_mass, self.cart_mass, self.length = pendulum_mass, cart_mass, length
        self.rot_friction, self.dt, self.gravity = rot_friction, dt, 9.81
        self.state_dim, self.action_dim = 4, 1
        self.normalization = normalization
        if normalization is not None:
            self.normalization = [np.array(n, dtype=np.float64) for n in normalization]
            self.inv_norm = [n ** -1 for n in self.normalization]
        self.input_dim, self.output_dim = 5, 4

    def linearize(self):
        m, M, L, b, g = (self.pendulum_mass, self.cart_mass, self.length, self.rot_friction,
                         self.gravity)
        A = np.array([[0, 0, 1, 0], [0, 0, 0, 1], [0, g * m / M, 0, -b / (M * L)],
                      [0, g * (m + M) / (L * M), 0, -b * (m + M) / (m * M * L ** 2)]],
                     dtype=np.float64)
        B = np.array([0, 0, 1 / M, 1 / (M * L)]).reshape((-1, 1))
        if self.normalization is not None:
            Tx, Tu = map(np.diag, self.normalization)
            Tx_inv, Tu_inv = map(np.diag, self.inv_norm)
            A = np.linalg.multi_dot((Tx_inv, A, Tx))
            B = np.linalg.multi_dot((Tx_inv, B, Tu))
        Ad, Bd, _, _, _ = scipy.signal.cont2discrete((A, B, 0, 0), self.dt, method="zoh")
        return Ad, Bd

    def descriptor(self):
        d = nat.SlbFunction()
        d.kind, d.in_dim, d.out_dim = nat.FN_CARTPOLE, 5, 4
        cp = d.cparams
        cp[0], cp[1], cp[2] = self.pendulum_mass, self.cart_mass, self.length
        cp[3], cp[4], cp[5] = self.rot_friction, self.gravity, self.dt / 10
        cp[15] = _pack_norm(cp, self.normalization, 4, 6)    # [6..9] Tx, [10] Tu, [11..14] 1/Tx
        return d


class LyapunovNetwork(DeterministicFunction):
    """Positive-definite network ``V(x) = |phi(x)|^2`` (``examples/utilities.py:48-104``),
    inference only: layer i applies ``act(net . [W_i^T W_i + eps I; W_i'']^T)``.

    ``weights[i] = (W_posdef, W_extra or None)`` with the reference's shapes
    (``[ceil((in+1)/2), in]`` and ``[out - in, in]``); when omitted they are drawn Xavier-uniform
    from ``seed`` (the reference uses ``tf.contrib.layers.xavier_initializer``).  ``activations``
    are 'tanh' | 'relu' | 'linear' (or ``numpy.tanh``).
    """

    _ACT = {"tanh": 0, "relu": 1, "linear": 2, "identity": 2}

    def __init__(self, input_dim, layer_dims, activations, eps=1e-6, initializer=None,
                 name="lyapunov_network", weights=None, seed=0):
        super().__init__(name)
        self.input_dim, self.output_dim = int(input_dim), 1
        self.num_layers = len(layer_dims)
        self.output_dims = [int(v) for v in layer_dims]
        self.eps = eps
        if self.output_dims[0] < self.input_dim:
            raise ValueError("The first layer dimension must be at least the input dimension!")
        if np.any(np.diff(self.output_dims) < 0):
            raise ValueError("Each layer must maintain or increase the dimension of its input!")
        if max(self.output_dims) > 64 or self.num_layers > 8:
            raise DimensionError("LyapunovNetwork: at most 8 layers of width <= 64 are fused")
        self.activations = []
        for act in activations:
            key = act if isinstance(act, str) else getattr(act, "__name__", "")
            if key not in self._ACT:
                raise NotImplementedError("activation %r is not fused (tanh/relu/linear)" % (act,))
            self.activations.append(key)
        self.hidden_dims = [int(np.ceil(((self.input_dim if i == 0 else self.output_dims[i - 1])
                                         + 1) / 2)) for i in range(self.num_layers)]
        if weights is None:
            rng = np.random.default_rng(seed)
            weights = []
            for i in range(self.num_layers):
                din = self.input_dim if i == 0 else self.output_dims[i - 1]
                def xavier(rows, cols):
                    lim = np.sqrt(6.0 / (rows + cols))
                    return rng.uniform(-lim, lim, size=(rows, cols))
                extra = self.output_dims[i] - din
                weights.append((xavier(self.hidden_dims[i], din),
                                xavier(extra, din) if extra > 0 else None))
        self.weights = weights
        self._kernel_dev = None

    def kernels(self):
        """Layer kernels ``[W^T W + eps I; W_extra]`` ([out_i, in_i])."""
        out = []
        for i, (w0, w1) in enumerate(self.weights):
            din = self.input_dim if i == 0 else self.output_dims[i - 1]
            k = w0.T.dot(w0) + self.eps * np.eye(din)
            if w1 is not None:
                k = np.concatenate([k, w1], axis=0)
            out.append(k)
        return out

    def descriptor(self):
        if self._kernel_dev is None:
            self._kernel_dev = dev.to_device(np.concatenate([k.ravel() for k in self.kernels()]))
        d = nat.SlbFunction()
        d.kind, d.in_dim, d.out_dim = nat.FN_LYAPUNOV_NN, self.input_dim, 1
        d.cparams[0] = self.num_layers
        for i, (od, act) in enumerate(zip(self.output_dims, self.activations)):
            d.cparams[1 + i] = od
            d.cparams[9 + i] = self._ACT[act]
        d.matrix = self._kernel_dev.data_ptr()
        return d


class NeuralNetwork(DeterministicFunction):
    """Dense MLP, inference only (``functions.py:1665-1729``): bias in the hidden layers only,
    no bias in the output layer, output multiplied by ``output_scale``.  ``layers`` =
    [in, h1, ..., out]; ``nonlinearities`` one per layer after the input ('tanh' | 'relu' | None).
    ``weights[i]`` has the TF layout ``[in_i, out_i]``, ``biases[i]`` ``[out_i]`` (hidden layers);
    both are drawn Xavier-uniform / zero from ``seed`` when omitted.  Training (the reference
    optimises these with TF optimisers) is outside this build.
    """

    def __init__(self, layers, nonlinearities, output_scale=1., use_bias=True,
                 name="neural_network", weights=None, biases=None, seed=0):
        super().__init__(name)
        self.layers = [int(v) for v in layers]
        self.nonlinearities = []
        for act in nonlinearities:
            key = "linear" if act is None else (act if isinstance(act, str)
                                                else getattr(act, "__name__", ""))
            if key not in LyapunovNetwork._ACT:
                raise NotImplementedError("nonlinearity %r is not fused (tanh/relu/None)" % (act,))
            self.nonlinearities.append(key)
        if len(self.nonlinearities) != len(self.layers) - 1:
            raise ValueError("one nonlinearity per layer after the input is required")
        self.output_scale = float(output_scale)
        self.use_bias = bool(use_bias)
        self.input_dim, self.output_dim = self.layers[0], self.layers[-1]
        if max(self.layers[1:]) > 64 or len(self.layers) - 1 > 8 or self.output_dim > nat.SLB_MAX_OUT:
            raise DimensionError("NeuralNetwork: at most 8 layers of width <= 64 are fused")
        rng = np.random.default_rng(seed)
        if weights is None:
            weights = []
            for din, dout in zip(self.layers[:-1], self.layers[1:]):
                lim = np.sqrt(6.0 / (din + dout))
                weights.append(rng.uniform(-lim, lim, size=(din, dout)))
        if biases is None:
            biases = [np.zeros(d) for d in self.layers[1:-1]]
        self.weights = [np.asarray(w, dtype=np.float64) for w in weights]
        self.biases = [np.asarray(b, dtype=np.float64) for b in biases]
        self._param_dev = None

    def descriptor(self):
        if self._param_dev is None:
            parts = []
            for i, w in enumerate(self.weights):
                parts.append(w.T.ravel())                       # [out, in] rows
                if self.use_bias and i + 1 < len(self.weights):
                    parts.append(self.biases[i].ravel())
            self._param_dev = dev.to_device(np.concatenate(parts))
        d = nat.SlbFunction()
        d.kind, d.in_dim, d.out_dim = nat.FN_MLP, self.input_dim, self.output_dim
        d.cparams[0] = len(self.weights)
        for i, (od, act) in enumerate(zip(self.layers[1:], self.nonlinearities)):
            d.cparams[1 + i] = od
            d.cparams[9 + i] = LyapunovNetwork._ACT[act]
        d.cparams[17] = self.output_scale
        d.cparams[18] = 1.0 if self.use_bias else 0.0
        d.matrix = self._param_dev.data_ptr()
        return d


# =============================================================================== Gaussian processes
class Kernel(object):
    """Covariance function with the algebra of ``gpflow==0.4.0`` ``kernels`` (third-party code the
    reference hands to ``GPRCached``, ``functions.py:370-393``): primitives carry ``input_dim`` and
    ``active_dims`` (default: the first ``input_dim`` columns), ``+`` and ``*`` build sums and
    products.  The device evaluates the sum-of-products normal form (``slb_kernel``)."""

    def __add__(self, other):
        return Sum([self, other])

    def __mul__(self, other):
        return Product([self, other])

    def terms(self):
        """Normal form: list of product terms, each a list of primitives."""
        raise NotImplementedError

    def hyper_key(self):
        return tuple(tuple(p.hyper_key() for p in term) for term in self.terms())

    def is_plain_rbf(self, din):
        return False

    def K_device(self, X, X2=None):
        """K(X, X2) on raw inputs (device tensors), gpflow arithmetic."""
        total = None
        for term in self.terms():
            prod = None
            for p in term:
                k = p._K(X, X2)
                prod = k if prod is None else prod * k
            total = prod if total is None else total + prod
        return total

    def Kdiag_device(self, X):
        total = None
        for term in self.terms():
            prod = None
            for p in term:
                k = p._Kdiag(X)
                prod = k if prod is None else prod * k
            total = prod if total is None else total + prod
        return total

    def fill(self, kstruct, din):
        """Write the normal form into an ``slb_kernel``."""
        terms = self.terms()
        count = sum(len(t) for t in terms)
        if count > nat.SLB_MAX_KPRIM:
            raise NotImplementedError("kernel expands to %d primitives; the device descriptor "
                                      "holds %d" % (count, nat.SLB_MAX_KPRIM))
        i = 0
        for t, term in enumerate(terms):
            for p in term:
                if max(p.active_dims) >= din:
                    raise DimensionError("kernel active_dims %r outside the %d GP input columns"
                                         % (p.active_dims, din))
                prim = kstruct.prims[i]
                prim.kind, prim.term, prim.variance = p.KIND, t, p._scalar_variance()
                weights = p._weights()
                for c in range(nat.SLB_MAX_IN):
                    prim.w[c] = 0.0
                for c, w in zip(p.active_dims, weights):
                    prim.w[c] = float(w)
                i += 1
        kstruct.num_prims = count


class Sum(Kernel):
    def __init__(self, kern_list):
        self.kern_list = list(kern_list)

    def terms(self):
        return [term for k in self.kern_list for term in k.terms()]


Add = Sum


class Product(Kernel):
    def __init__(self, kern_list):
        self.kern_list = list(kern_list)

    def terms(self):
        out = [[]]
        for k in self.kern_list:
            out = [a + b for a in out for b in k.terms()]
        return out


Prod = Product


class _Primitive(Kernel):
    KIND = None

    def __init__(self, input_dim, active_dims=None):
        self.input_dim = int(input_dim)
        if active_dims is None:
            active_dims = range(self.input_dim)
        elif isinstance(active_dims, slice):
            active_dims = range(*active_dims.indices(1 << 30))[:self.input_dim]
        self.active_dims = [int(a) for a in active_dims]
        if len(self.active_dims) != self.input_dim:
            raise DimensionError("active_dims %r does not select input_dim = %d columns"
                                 % (self.active_dims, self.input_dim))

    def terms(self):
        return [[self]]

    def _slice(self, X):
        return X[:, self.active_dims]

    def _scalar_variance(self):
        return float(self.variance)

    def _weights(self):
        return np.zeros(self.input_dim)

    def hyper_key(self):
        return (self.KIND, tuple(self.active_dims), self._scalar_variance(),
                tuple(np.asarray(self._weights(), dtype=np.float64).tolist()))


class _Stationary(_Primitive):
    def __init__(self, input_dim, variance=1.0, lengthscales=None, active_dims=None, ARD=False):
        _Primitive.__init__(self, input_dim, active_dims)
        self.variance = float(variance)
        ls = 1.0 if lengthscales is None else lengthscales
        self.lengthscales = np.broadcast_to(np.asarray(ls, dtype=np.float64),
                                            (self.input_dim,)).copy()
        self.ARD = ARD

    def _weights(self):
        return 1.0 / self.lengthscales

    def hyper_key(self):
        return (self.KIND, tuple(self.active_dims), self.variance,
                tuple(self.lengthscales.tolist()))

    def _square_dist(self, X, X2):
        """gpflow ``Stationary.square_dist``: the |x|^2 + |x'|^2 - 2 x.x' expansion."""
        ls = torch.as_tensor(self.lengthscales, dtype=torch.float64, device=X.device)
        X = self._slice(X) / ls
        sq = (X * X).sum(dim=1)
        if X2 is None:
            return -2.0 * (X @ X.T) + sq[:, None] + sq[None, :]
        X2 = self._slice(X2) / ls
        sq2 = (X2 * X2).sum(dim=1)
        return -2.0 * (X @ X2.T) + sq[:, None] + sq2[None, :]

    def _euclid_dist(self, X, X2):
        return torch.sqrt(self._square_dist(X, X2) + 1e-12)

    def _Kdiag(self, X):
        return torch.full((X.shape[0],), self.variance, dtype=torch.float64, device=X.device)


class RBF(_Stationary):
    """Squared-exponential kernel with the arithmetic of ``gpflow==0.4.0`` ``kernels.RBF``:
    ``variance * exp(-0.5 * sum(((x - x') / lengthscales)^2))``, ``Kdiag = variance``,
    defaults 1 (``lengthscales`` scalar or ARD vector)."""
    KIND = nat.K_RBF

    def is_plain_rbf(self, din):
        return self.active_dims == list(range(din))

    def _K(self, X, X2=None):
        return self.variance * torch.exp(-self._square_dist(X, X2) / 2)

    def K_scaled(self, Xs):
        """K(X, X) from lengthscale-divided inputs (device tensor [M, d])."""
        sq = (Xs * Xs).sum(dim=1)
        dist = -2.0 * (Xs @ Xs.T) + sq[:, None] + sq[None, :]
        return self.variance * torch.exp(-dist / 2)


class Matern12(_Stationary):
    """``variance * exp(-r)``, ``r = sqrt(square_dist + 1e-12)`` (gpflow 0.4.0)."""
    KIND = nat.K_MATERN12

    def _K(self, X, X2=None):
        return self.variance * torch.exp(-self._euclid_dist(X, X2))


class Matern32(_Stationary):
    """``variance * (1 + sqrt(3) r) * exp(-sqrt(3) r)`` (gpflow 0.4.0)."""
    KIND = nat.K_MATERN32

    def _K(self, X, X2=None):
        r = self._euclid_dist(X, X2)
        return self.variance * (1. + np.sqrt(3.) * r) * torch.exp(-np.sqrt(3.) * r)


class Matern52(_Stationary):
    """``variance * (1 + sqrt(5) r + 5/3 r^2) * exp(-sqrt(5) r)`` (gpflow 0.4.0)."""
    KIND = nat.K_MATERN52

    def _K(self, X, X2=None):
        r = self._euclid_dist(X, X2)
        return self.variance * (1. + np.sqrt(5.) * r + 5. / 3. * r * r) * torch.exp(-np.sqrt(5.) * r)


class Linear(_Primitive):
    """gpflow 0.4.0 ``kernels.Linear``: ``K = (X * variance) X'^T`` on the active columns,
    ``variance`` a scalar or (``ARD=True``) one value per column."""
    KIND = nat.K_LINEAR

    def __init__(self, input_dim, variance=1.0, active_dims=None, ARD=False):
        _Primitive.__init__(self, input_dim, active_dims)
        self.ARD = ARD
        self.variance = np.broadcast_to(np.asarray(variance, dtype=np.float64),
                                        (self.input_dim,)).copy()

    def _scalar_variance(self):
        return 1.0

    def _weights(self):
        return self.variance

    def _K(self, X, X2=None):
        var = torch.as_tensor(self.variance, dtype=torch.float64, device=X.device)
        X = self._slice(X)
        X2 = X if X2 is None else self._slice(X2)
        return (X * var) @ X2.T

    def _Kdiag(self, X):
        var = torch.as_tensor(self.variance, dtype=torch.float64, device=X.device)
        X = self._slice(X)
        return (X * X * var).sum(dim=1)


class Constant(_Primitive):
    """gpflow 0.4.0 ``kernels.Constant`` / ``Bias``: ``K = variance`` everywhere."""
    KIND = nat.K_CONSTANT

    def __init__(self, input_dim, variance=1.0, active_dims=None):
        _Primitive.__init__(self, input_dim, active_dims)
        self.variance = float(variance)

    def _K(self, X, X2=None):
        n2 = X.shape[0] if X2 is None else X2.shape[0]
        return torch.full((X.shape[0], n2), self.variance, dtype=torch.float64, device=X.device)

    def _Kdiag(self, X):
        return torch.full((X.shape[0],), self.variance, dtype=torch.float64, device=X.device)


Bias = Constant


class White(_Primitive):
    """gpflow 0.4.0 ``kernels.White``: ``variance * I`` for ``K(X)``, zeros against new points."""
    KIND = nat.K_WHITE

    def __init__(self, input_dim, variance=1.0, active_dims=None):
        _Primitive.__init__(self, input_dim, active_dims)
        self.variance = float(variance)

    def _K(self, X, X2=None):
        if X2 is None:
            return self.variance * torch.eye(X.shape[0], dtype=torch.float64, device=X.device)
        return torch.zeros((X.shape[0], X2.shape[0]), dtype=torch.float64, device=X.device)

    def _Kdiag(self, X):
        return torch.full((X.shape[0],), self.variance, dtype=torch.float64, device=X.device)


class _KernelNamespace(object):
    """``gpflow.kernels``-style access: ``safe_learning_b200.kernels.Matern32(...)``."""
    RBF, Matern12, Matern32, Matern52 = RBF, Matern12, Matern32, Matern52
    Linear, Constant, Bias, White, Add, Prod = Linear, Constant, Bias, White, Add, Prod


kernels = _KernelNamespace()


class Likelihood(object):
    """Gaussian likelihood holder (``gp.likelihood.variance``, default 1 like gpflow)."""

    def __init__(self, variance=1.0):
        self.variance = float(variance)


class _Factor(object):
    """Device-resident Cholesky state of one (X, kernel, noise, scale) combination."""

    __slots__ = ("key", "M", "nrb", "Xs", "L", "Linv", "Wpack", "Whead", "Wheadp", "Xhead", "head_rows",
                 "Xf", "hmax", "appends", "plain", "floor_rel")


_FACTOR_CACHE = {}
_FACTOR_CACHE_MAX = 8


def _remember_factor(fac):
    if len(_FACTOR_CACHE) >= _FACTOR_CACHE_MAX:
        _FACTOR_CACHE.pop(next(iter(_FACTOR_CACHE)))
    _FACTOR_CACHE[fac.key] = fac


class GPRCached(object):
    """GP regression with the factor cached in HBM (``functions.py:357-458``).

    ``update_cache`` (``:395-415``) builds ``L = chol(scale^2 (K + noise I))`` with torch's
    cuSOLVER/cuBLAS on the device (library plumbing, between sweeps), forms ``L^-1`` and packs
    it into the fragment order the sweep kernel streams (``slb_pack_factor``), and
    ``alpha = L^-1 scale (Y - m(X))``.  GPs that share X, kernel, noise and scale share one
    factor (the stacked GPs of a ``FunctionStack`` often do).
    """

    def __init__(self, x, y, kern, mean_function=None, scale=1., name="GPRCached",
                 noise_variance=1.0):
        self.name = name
        self._X = np.atleast_2d(np.asarray(x, dtype=np.float64))
        self._Y = np.atleast_2d(np.asarray(y, dtype=np.float64))
        if self._Y.shape[1] != 1:
            raise DimensionError("one-output GPs only; stack them with FunctionStack")
        if not isinstance(kern, Kernel):
            raise TypeError("kern must be built from safe_learning_b200 kernels (RBF, Matern12/32/52, "
                            "Linear, Constant, White and their sums / products), got %r"
                            % type(kern).__name__)
        if mean_function is not None and not isinstance(mean_function, LinearSystem):
            raise NotImplementedError("prior mean must be None or a one-row LinearSystem")
        if self._X.shape[1] > nat.SLB_MAX_IN:
            raise DimensionError("GP input dimension above %d" % nat.SLB_MAX_IN)
        self.kern = kern
        self.mean_function = mean_function
        self.likelihood = Likelihood(noise_variance)
        self._scale = float(scale)
        self._factor = None
        self._alpha_dev = None
        self._gamma_dev = None
        self._gamma_f_dev = None
        self._gamma_l1 = 0.0
        self._prior_dev = None
        self._stale = True
        self._version = 0
        self._hyper_seen = None

    # data ------------------------------------------------------------------------------
    @property
    def X(self):
        return self._X

    @X.setter
    def X(self, value):
        self._X = np.atleast_2d(np.asarray(value, dtype=np.float64))
        self._stale = True

    @property
    def Y(self):
        return self._Y

    @Y.setter
    def Y(self, value):
        self._Y = np.atleast_2d(np.asarray(value, dtype=np.float64))
        self._stale = True

    @property
    def cholesky(self):
        self._ensure()
        return self._factor.L.cpu().numpy()

    @property
    def alpha(self):
        self._ensure()
        return self._alpha_dev[:self._X.shape[0]].cpu().numpy()[:, None]

    # factorisation -----------------------------------------------------------------------
    def _factor_key(self):
        digest = hashlib.sha1(np.ascontiguousarray(self._X).tobytes()).hexdigest()
        return (digest, self._X.shape, self.kern.hyper_key(), self.likelihood.variance,
                self._scale)

    def _hyper_state(self):
        return (self.kern.hyper_key(), self.likelihood.variance, self._scale)

    def _ensure(self):
        """Refresh the factor if data (setters mark it stale) or hyper-parameters changed."""
        if self._stale or self._factor is None or self._hyper_seen != self._hyper_state():
            self.update_cache()

    @property
    def version(self):
        self._ensure()
        return self._version

    def update_cache(self):
        """``functions.py:395-415`` on the device."""
        lib = nat.load()
        key = self._factor_key()
        M, din = self._X.shape
        fac = _FACTOR_CACHE.get(key)
        if fac is None:
            fac = _Factor()
            fac.key, fac.M, fac.nrb = key, M, (M + 7) // 8
            fac.plain = self.kern.is_plain_rbf(din)
            fac.appends = 0
            if M == 0:
                # empty data set (the notebooks start from np.empty((0, d))): prior only
                fac.Xs = dev.zeros((0, din))
                fac.L = fac.Linv = dev.zeros((0, 0))
                fac.Wpack = dev.zeros((1,))
                self._head(fac, None)
            else:
                if fac.plain:
                    fac.Xs = dev.to_device(self._X / self.kern.lengthscales)
                    kernel = self.kern.K_scaled(fac.Xs)
                else:
                    fac.Xs = dev.to_device(self._X)
                    kernel = self.kern.K_device(fac.Xs)
                kernel_noisy = kernel + torch.eye(M, dtype=torch.float64, device=kernel.device) \
                    * self.likelihood.variance
                kernel_noisy = kernel_noisy * (self._scale ** 2)
                fac.L = torch.linalg.cholesky(kernel_noisy)
                fac.Linv = torch.linalg.solve_triangular(
                    fac.L, torch.eye(M, dtype=torch.float64, device=kernel.device),
                    upper=False).contiguous()
                self._pack(fac, kernel)
            _remember_factor(fac)
        self._finish_cache(fac)

    def _pack(self, fac, kernel=None, head_from=None):
        """Device tables derived from ``L^-1``: the DMMA-ordered packed factor and the tables of
        the decision filter (``_head``).  ``kernel``: ``K(X, X)`` without noise if the caller has
        it; ``head_from``: an older factor of the same model whose head subset is kept."""
        lib = nat.load()
        fac.Wpack = dev.empty((int(lib.slb_packed_len(fac.M)),))
        nat.check(lib.slb_pack_factor(dev.stream(), fac.Linv.data_ptr(), fac.M,
                                      fac.Wpack.data_ptr()), "slb_pack_factor")
        self._head(fac, kernel, head_from)

    def _head(self, fac, kernel, head_from=None):
        """Tables of the decision filter (``slb_lyapunov_sweep_filtered``, csrc/filter.cu).

        * Head subset: the posterior variance given ANY subset S of the training set bounds the
          full posterior variance from above.  S = the first ``min(M, SLB_HEAD_RANK)`` points in
          pivoted-Cholesky order of ``K(X, X)`` (greedy: the point with the largest variance given
          the ones already chosen), factored on its own: ``Whead = chol(scale^2 (K_SS + noise
          I))^-1`` (stored transposed = column-major, zero padded), ``Xhead = Xs[S]``.  A factor
          grown by ``append_data`` keeps the subset of the factor it grew from.
        * ``Xf``: the training inputs as the filter's mean stage streams them (TMA bulk copies:
          rows padded to a multiple of 4, 16-byte aligned); for the plain RBF each row carries
          ``-|x / l|^2 / 2`` so that the squared distance expands into three FMAs.
        * ``floor_rel``: certified lower bound of posterior variance / prior variance that
          decides whether the filter may be used at all (``Lyapunov._filter_enabled``)."""
        R = nat.SLB_HEAD_RANK
        M, din = fac.M, self._X.shape[1]
        r = min(M, R)
        fac.Whead = dev.zeros((R, R))
        fac.Xhead = dev.zeros((R, din))
        fac.Wheadp = dev.zeros((R * R,))
        fac.head_rows = r
        Mp = max(8 * ((M + 7) // 8), 8)
        width = din + 1 if fac.plain else din
        fac.Xf = dev.zeros((Mp, width))
        fac.hmax = 0.0
        if M == 0:
            fac.floor_rel = 1.0
            return
        fac.Xf[:M, :din] = fac.Xs
        if fac.plain:
            half = -0.5 * (fac.Xs * fac.Xs).sum(dim=1)
            fac.Xf[:M, din] = half
            fac.hmax = float(-half.min().item())
        fac.Wheadp = dev.zeros((R * R,))
        if head_from is not None and head_from.head_rows == r:
            fac.Whead, fac.Wheadp, fac.Xhead = head_from.Whead, head_from.Wheadp, head_from.Xhead
        else:
            if kernel is None:
                kernel = self.kern.K_scaled(fac.Xs) if fac.plain else self.kern.K_device(fac.Xs)
            subset = self._pivoted_subset(kernel, r)
            fac.Xhead[:r] = fac.Xs.index_select(0, subset)
            k_ss = kernel.index_select(0, subset).index_select(1, subset)
            k_ss = (k_ss + torch.eye(r, dtype=torch.float64, device=k_ss.device)
                    * self.likelihood.variance) * (self._scale ** 2)
            l_ss = torch.linalg.cholesky(k_ss)
            linv = torch.linalg.solve_triangular(
                l_ss, torch.eye(r, dtype=torch.float64, device=k_ss.device), upper=False)
            fac.Whead[:r, :r] = linv.T                    # Whead[j, i] = L_S^-1[i, j]
            # the same matrix in DMMA A-fragment order (head stage of the filter): [b, s, T/4, T%4]
            full = dev.zeros((R, R))
            full[:r, :r] = linv
            fac.Wheadp = full.reshape(R // 8, 8, R // 4, 4).permute(0, 2, 1, 3).contiguous().reshape(-1)
        # var(z) >= k(z,z) s / (M k(z,z) + s), s = noise variance (M noisy observations AT z are
        # the most informative data set): relative to k(z,z) at least s / (M kmax + s).  When
        # that is not far above fp64 rounding of the O(M^2) contraction the reference's
        # "negative variance -> NaN -> unsafe" corner (functions.py:451) is reachable and the
        # filter, which bounds the variance from above only, is not used.  Neither is it with
        # non-finite or absurdly large inputs (the expanded distance needs moderate magnitudes).
        noise = float(self.likelihood.variance)
        if fac.plain:
            kmax = float(self.kern.variance)
        else:
            kmax = float(self.kern.Kdiag_device(fac.Xs).max().item())
        fac.floor_rel = noise / (fac.M * kmax + noise) if (kmax > 0 or noise > 0) else 0.0
        if not bool(torch.isfinite(fac.Xs).all()) or float(fac.Xs.abs().max().item()) > 1e100:
            fac.floor_rel = 0.0

    @staticmethod
    def _pivoted_subset(kernel, r):
        """Indices of the first ``r`` pivots of the pivoted Cholesky factorisation of the symmetric
        positive semi-definite ``kernel`` (device tensor [M, M]) -> int64 device tensor [r]: one
        kernel launch (``slb_pivoted_subset``), no host synchronisation."""
        lib = nat.load()
        M = kernel.shape[0]
        kernel = kernel.contiguous()
        picks = dev.zeros((r,), torch.int64)
        scratch = dev.empty((M * (r + 1),))
        nat.check(lib.slb_pivoted_subset(dev.stream(), kernel.data_ptr(), M, r, picks.data_ptr(),
                                         scratch.data_ptr()), "slb_pivoted_subset")
        return picks

    def _append_rows(self, x_new):
        """Rank-one growth of the cached factor for each appended observation (SURVEY.md 8f
        item 2) -- O(M^2) on the device instead of the O(M^3) refactorisation of
        ``functions.py:395-415``:  L' = [[L, 0], [l^T, lam]],  l = L^-1 k,  lam = sqrt(k** - l.l),
        L'^-1 = [[L^-1, 0], [-(l^T L^-1) / lam, 1 / lam]].  Returns the new factor, or None when
        a full refit is due (every 256 appends, or if the pivot is not safely positive)."""
        old = self._factor
        if old is None or self._hyper_seen != self._hyper_state():
            return None
        key = self._factor_key()
        cached = _FACTOR_CACHE.get(key)
        if cached is not None:
            return cached
        x_new = np.atleast_2d(np.asarray(x_new, dtype=np.float64))
        if old.appends + len(x_new) > 256 or old.M + len(x_new) != self._X.shape[0]:
            return None
        s2 = self._scale ** 2
        Xs, L, Linv = old.Xs, old.L, old.Linv
        if old.M == 0:
            return None
        for row in x_new:
            M = Xs.shape[0]
            if old.plain:
                xs = dev.to_device((row / self.kern.lengthscales)[None, :])
                # same expansion as RBF.K_scaled (gpflow square_dist)
                dist = -2.0 * (Xs @ xs.T)[:, 0] + (Xs * Xs).sum(dim=1) + (xs * xs).sum()
                k = s2 * (self.kern.variance * torch.exp(-dist / 2))
                kss = s2 * (self.kern.variance + self.likelihood.variance)
            else:
                xs = dev.to_device(row[None, :])
                k = s2 * self.kern.K_device(Xs, xs)[:, 0]
                kss = s2 * (self.kern.Kdiag_device(xs)[0] + self.likelihood.variance)
            l = Linv @ k
            lam2 = kss - torch.dot(l, l)
            if not bool(lam2 > 1e-12 * kss):
                return None
            lam = torch.sqrt(lam2)
            L_new = torch.zeros((M + 1, M + 1), dtype=torch.float64, device=L.device)
            L_new[:M, :M] = L
            L_new[M, :M] = l
            L_new[M, M] = lam
            Linv_new = torch.zeros_like(L_new)
            Linv_new[:M, :M] = Linv
            Linv_new[M, :M] = -(l @ Linv) / lam
            Linv_new[M, M] = 1.0 / lam
            Xs, L, Linv = torch.cat((Xs, xs), dim=0), L_new, Linv_new
        fac = _Factor()
        fac.key, fac.M, fac.nrb = key, Xs.shape[0], (Xs.shape[0] + 7) // 8
        fac.Xs, fac.L, fac.Linv = Xs.contiguous(), L, Linv.contiguous()
        fac.appends, fac.plain = old.appends + len(x_new), old.plain
        self._pack(fac, None, head_from=old if old.head_rows == nat.SLB_HEAD_RANK else None)
        _remember_factor(fac)
        return fac

    def append_data(self, x, y):
        """Append observations; grows the cached factor incrementally when possible."""
        self._X = np.vstack((self._X, np.atleast_2d(x)))
        self._Y = np.vstack((self._Y, np.atleast_2d(y)))
        was_fresh = not self._stale
        self._stale = True
        fac = self._append_rows(x) if was_fresh else None
        if fac is None:
            self.update_cache()
        else:
            self._finish_cache(fac)

    def _finish_cache(self, fac):
        """alpha / gamma / prior mean for this GP's targets on factor `fac` (functions.py:405-409)."""
        M = fac.M
        self._factor = fac
        target = dev.to_device(self._Y)
        if self.mean_function is not None:
            if M:
                target = target - self.mean_function.evaluate_device(self._X)
            self._prior_dev = dev.to_device(self.mean_function.matrix.reshape(-1))
        else:
            self._prior_dev = None
        target = self._scale * target
        alpha = fac.Linv @ target
        gamma = fac.Linv.T @ alpha
        padded = dev.zeros((max(8 * fac.nrb, 8),))
        padded[:M] = alpha[:, 0]
        self._alpha_dev = padded
        self._gamma_dev = gamma[:, 0].contiguous() if M else dev.zeros((1,))
        # the filter's mean weights: scale^2 gamma, times the RBF variance on the plain path
        # (whose kernel values are then <= 1), zero padded to the row count of Xf; and the bound
        # sum_i (|L^-1|^T |alpha|)_i of everything that rounds when the mean is summed
        fold = self._scale ** 2 * (float(self.kern.variance) if fac.plain else 1.0)
        self._gamma_f_dev = dev.zeros((fac.Xf.shape[0],))
        self._gamma_l1 = 0.0
        if M:
            self._gamma_f_dev[:M] = fold * gamma[:, 0]
            self._gamma_l1 = abs(fold) * float((fac.Linv.abs().T @ alpha.abs()).sum().item())
        self._stale = False
        self._hyper_seen = self._hyper_state()
        self._version += 1

    # differentiable prediction -------------------------------------------------------------
    def torch_predict(self, points):
        """``build_predict`` (``functions.py:417-458``) on a device tensor [n, d_in] in torch
        operations (cuBLAS / cuSOLVER on the cached factor), differentiable with respect to
        ``points``: latent mean and variance, [n, 1] each."""
        self._ensure()
        fac, s = self._factor, self._scale
        s2 = s * s
        mx = 0.0
        if self.mean_function is not None:
            self.mean_function.descriptor()
            mx = s * (points @ self.mean_function._matrix_dev.T)               # :439
        kss = s2 * self.kern.Kdiag_device(points)                                # :450
        if fac.M == 0:
            return mx / s + 0.0 * points[:, :1], (kss / s2).unsqueeze(1)
        x_train = dev.to_device(self._X)
        kx = s2 * self.kern.K_device(x_train, points)                            # :438  [M, n]
        a = torch.linalg.solve_triangular(fac.L, kx, upper=False)                # :441
        alpha = self._alpha_dev[:fac.M].unsqueeze(1)
        fmean = (a.T @ alpha + mx) / s                                           # :442, :455
        fvar = (kss - (a * a).sum(dim=0)) / s2                                   # :451, :456
        return fmean, fvar.unsqueeze(1)

    # descriptor pieces -------------------------------------------------------------------
    def fill_factor(self, f):
        self._ensure()
        fac = self._factor
        f.M, f.nrb = fac.M, fac.nrb
        f.Xs, f.Wpack = fac.Xs.data_ptr(), fac.Wpack.data_ptr()
        f.Whead, f.Xhead, f.head_rows = fac.Whead.data_ptr(), fac.Xhead.data_ptr(), fac.head_rows
        f.Wheadp = fac.Wheadp.data_ptr()
        f.Xf, f.hmax = fac.Xf.data_ptr(), fac.hmax
        f.scale = self._scale
        if fac.plain:
            for c, ls in enumerate(self.kern.lengthscales):
                f.lengthscales[c] = float(ls)
            f.variance = self.kern.variance
            f.kss = (self._scale ** 2) * self.kern.variance
            f.kernel.num_prims = 0
        else:
            self.kern.fill(f.kernel, self._X.shape[1])
        return fac.key

    def fill_output(self, o, factor_index, beta):
        self._ensure()
        o.factor, o.beta = factor_index, float(beta)
        o.alpha = self._alpha_dev.data_ptr()
        o.gamma = self._gamma_dev.data_ptr()
        o.prior_mean = None if self._prior_dev is None else self._prior_dev.data_ptr()
        o.gamma_f, o.gamma_l1 = self._gamma_f_dev.data_ptr(), self._gamma_l1


    def variance_floor(self):
        """Certified lower bound of posterior variance / prior variance (see ``_head``)."""
        self._ensure()
        return self._factor.floor_rel

    # host copies of the cached tables ------------------------------------------------------
    _CACHE_FIELDS = ("Xs", "Wpack", "Whead", "Wheadp", "Xhead", "Xf", "alpha", "gamma", "gamma_f")

    def _cache_tables(self):
        fac = self._factor
        return (fac.Xs, fac.Wpack, fac.Whead, fac.Wheadp, fac.Xhead, fac.Xf, self._alpha_dev,
                self._gamma_dev, self._gamma_f_dev)

    def export_cache(self, pinned=True):
        """Host copies (torch CPU tensors, page-locked if ``pinned``) of the device tables a sweep
        reads: scaled training inputs, packed ``L^-1``, its head block, ``alpha`` and ``gamma``
        (what ``update_cache`` leaves in HBM, ``functions.py:395-415``).  Together with
        ``import_cache`` this checkpoints / restores the cached state without a refit."""
        self._ensure()
        fac = self._factor
        out = {}
        for name, t in zip(self._CACHE_FIELDS, self._cache_tables()):
            host = t.detach().cpu()
            out[name] = host.pin_memory() if pinned and torch.cuda.is_available() else host
        return out

    def import_cache(self, tables):
        """Copy tables produced by ``export_cache`` (same data set and hyper-parameters, hence same
        shapes) back into the device buffers: asynchronous H2D copies on the current stream."""
        self._ensure()
        fac = self._factor
        for name, dst in zip(self._CACHE_FIELDS, self._cache_tables()):
            src = tables[name]
            if not isinstance(src, torch.Tensor):
                src = torch.from_numpy(np.ascontiguousarray(src, dtype=np.float64))
            if tuple(src.shape) != tuple(dst.shape):
                raise DimensionError("import_cache: %s has shape %s, the cached table %s"
                                     % (name, tuple(src.shape), tuple(dst.shape)))
            dst.copy_(src, non_blocking=True)
        return sum(int(tables[k].numel()) * 8 for k in self._CACHE_FIELDS)


GPR = GPRCached     # the uncached gpflow.gpr.GPR of the notebooks maps onto the cached one

def _table_slots(gps):
    """(owner, attribute) of every cached device table of a list of GPRCached models, each table
    once (stacked GPs may share a factor)."""
    slots, seen = [], set()
    for gp in gps:
        gp._ensure()
        fac = gp._factor
        for owner, names in ((fac, ("Xs", "Wpack", "Whead", "Wheadp", "Xhead", "Xf")),
                             (gp, ("_alpha_dev", "_gamma_dev", "_gamma_f_dev"))):
            for name in names:
                if (id(owner), name) not in seen:
                    seen.add((id(owner), name))
                    slots.append((owner, name))
    return slots


class PackedCache(object):
    """All cached GP tables of a model (or stack) in ONE contiguous device arena with a page-locked
    host mirror, so that checkpoint / restore is a single copy each way.  Building it re-homes the
    tables into the arena (their device pointers change: descriptors are rebuilt)."""

    ALIGN = 32                     # doubles: 256-byte table alignment (TMA bulk copies need 16 bytes)

    def __init__(self, gps, pinned=True):
        self.gps = list(gps)
        # the packed factors L^-1 (90% of the bytes; read by the O(M^2) posterior only) go last, so
        # that restore() can send them on a second stream behind the tables the filter stages read
        slots = _table_slots(self.gps)
        self.slots = ([sl for sl in slots if sl[1] != "Wpack"] + [sl for sl in slots if sl[1] == "Wpack"])
        offsets, total = [], 0
        self.split = None
        for owner, name in self.slots:
            if name == "Wpack" and self.split is None:
                self.split = total
            offsets.append(total)
            total += -(-max(getattr(owner, name).numel(), 1) // self.ALIGN) * self.ALIGN
        if self.split is None:
            self.split = total
        self.arena = dev.zeros((total,))
        self.views = []
        for (owner, name), off in zip(self.slots, offsets):
            old = getattr(owner, name)
            view = self.arena[off:off + old.numel()].view(old.shape)
            view.copy_(old)
            setattr(owner, name, view)
            self.views.append(view)
        for gp in self.gps:
            gp._version += 1       # the descriptors point at the old buffers
        host = self.arena.cpu()
        self.host = host.pin_memory() if pinned and torch.cuda.is_available() else host
        self.nbytes = int(total) * 8
        self.token = tuple(id(gp._factor) for gp in self.gps)
        self._side = None
        self._pinned = bool(self.host.is_pinned())

    def valid(self):
        return (self.token == tuple(id(gp._factor) for gp in self.gps)
                and all(getattr(o, n) is v for (o, n), v in zip(self.slots, self.views)))

    def restore(self, overlap=True):
        """Host mirror -> device arena, asynchronously.  With ``overlap`` (default, page-locked
        mirror) the small tables travel on the current stream and the packed factors on a second
        one; every launch that reads the packed factors waits for that copy inside the library
        (``slb_record_factor_dependency``), so a sweep enqueued right after this call runs its filter
        stages while the factors are still arriving.  Returns the bytes copied."""
        total = self.arena.numel()
        lib = nat.load()
        cur = torch.cuda.current_stream().cuda_stream
        if not (overlap and self._pinned and 0 < self.split < total):
            nat.check(lib.slb_restore_tables(self.arena.data_ptr(), self.host.data_ptr(), 8 * total, 8 * total,
                                             cur, None), "slb_restore_tables")
            return self.nbytes
        if self._side is None:
            self._side = torch.cuda.Stream()
        nat.check(lib.slb_restore_tables(self.arena.data_ptr(), self.host.data_ptr(), 8 * self.split,
                                         8 * total, cur, self._side.cuda_stream), "slb_restore_tables")
        dev.note_factor_dependency()
        return self.nbytes




def _build_stack(gps, betas):
    """slb_gp_stack for a list of GPRCached models (factor sharing by key)."""
    stack = nat.SlbGpStack()
    if len(gps) > nat.SLB_MAX_OUT:
        raise DimensionError("at most %d stacked GPs" % nat.SLB_MAX_OUT)
    stack.num_outputs = len(gps)
    stack.input_dim = gps[0].X.shape[1]
    keys = []
    for o, (gp, beta) in enumerate(zip(gps, betas)):
        if gp.X.shape[1] != stack.input_dim:
            raise DimensionError("stacked GPs must share the input dimension")
        gp._ensure()
        key = gp._factor.key
        if key not in keys:
            gp.fill_factor(stack.factors[len(keys)])
            keys.append(key)
        gp.fill_output(stack.outputs[o], keys.index(key), beta)
    stack.num_factors = len(keys)
    return stack


def _gp_predict(stack, points, want_var=False):
    lib = nat.load()
    pts = dev.to_device(points)
    if pts.dim() != 2 or pts.shape[1] != stack.input_dim:
        raise DimensionError("GP expects %d input columns, got %s"
                             % (stack.input_dim, tuple(pts.shape)))
    n, D = pts.shape[0], stack.num_outputs
    mean, err = dev.empty((n, D)), dev.empty((n, D))
    nat.check(lib.slb_gp_predict(dev.stream(), stack, pts.data_ptr(), n, mean.data_ptr(),
                                 err.data_ptr(), 1 if want_var else 0), "slb_gp_predict")
    return mean, err


class GaussianProcess(UncertainFunction):
    """``(mean, beta * sqrt(var))`` of a one-output GP (``functions.py:461-546``)."""

    def __init__(self, gaussian_process, beta=2., name="gaussian_process"):
        super().__init__(name)
        if not isinstance(gaussian_process, GPRCached):
            raise TypeError("gaussian_process must be a safe_learning_b200 GPRCached/GPR model")
        self.gaussian_process = gaussian_process
        self.beta = float(beta)
        self.n_dim = self.input_dim = gaussian_process.X.shape[1]
        self.output_dim = gaussian_process.Y.shape[1]

    @property
    def X(self):
        return self.gaussian_process.X

    @property
    def Y(self):
        return self.gaussian_process.Y

    def gp_stack(self):
        return _build_stack([self.gaussian_process], [self.beta])

    def variance_floor(self):
        return self.gaussian_process.variance_floor()

    def export_cache(self, pinned=True):
        """See ``FunctionStack.export_cache``."""
        return PackedCache([self.gaussian_process], pinned)

    def import_cache(self, tables):
        return FunctionStack.import_cache(self, tables)

    @property
    def version(self):
        return (id(self.gaussian_process), self.gaussian_process.version, self.beta)

    def __call__(self, *inputs):
        mean, err = _gp_predict(self.gp_stack(), concatenate_inputs(inputs))
        return mean.cpu().numpy(), err.cpu().numpy()

    def predict_device(self, points, want_var=False):
        return _gp_predict(self.gp_stack(), points, want_var)

    def torch(self, points):
        """(mean, beta * sigma) as differentiable torch tensors [n, 1] (``functions.py:507-515``)."""
        mean, var = self.gaussian_process.torch_predict(points)
        return mean, self.beta * torch.sqrt(var)

    def update_feed_dict(self):
        """Reference hook (``functions.py:517-523``): hyper-parameters travel in the
        descriptor here, so this only refreshes the cached factor."""
        self.gaussian_process._ensure()

    def add_data_point(self, x, y):
        """Append observations and refresh the factor (``functions.py:525-546``)."""
        self.gaussian_process.append_data(x, y)


class FunctionStack(UncertainFunction):
    """Stack of one-output GPs, one per state dimension (``functions.py:254-307``)."""

    def __init__(self, functions, name="function_stack"):
        super().__init__(name)
        self.functions = list(functions)
        for f in self.functions:
            if not isinstance(f, GaussianProcess):
                raise TypeError("FunctionStack fuses GaussianProcess members only")
        self.num_fun = len(self.functions)
        self.input_dim = self.functions[0].input_dim
        self.output_dim = sum(f.output_dim for f in self.functions)

    def gp_stack(self):
        return _build_stack([f.gaussian_process for f in self.functions],
                            [f.beta for f in self.functions])

    def variance_floor(self):
        return min(f.variance_floor() for f in self.functions)

    def export_cache(self, pinned=True):
        """Checkpoint of everything a sweep reads from the cached GPs (scaled training inputs,
        packed ``L^-1``, filter tables, ``alpha``, ``gamma`` -- what ``update_cache`` leaves in HBM,
        ``functions.py:395-415``) as a ``PackedCache``: the tables are re-homed into one contiguous
        device arena mirrored by one page-locked host buffer, so ``import_cache`` is a single H2D
        copy.  (``GPRCached.export_cache`` gives per-table host tensors instead.)"""
        return PackedCache([f.gaussian_process for f in self.functions], pinned)

    def import_cache(self, tables):
        """Restore the cached tables from ``export_cache``'s result (same data set and hyper-
        parameters); returns the bytes copied host -> device."""
        if isinstance(tables, PackedCache):
            if not tables.valid():
                raise DimensionError("import_cache: the GP cache was refitted since export_cache")
            return tables.restore()
        gps = [f.gaussian_process for f in getattr(self, "functions", [self])]
        return sum(gp.import_cache(t) for gp, t in zip(gps, tables))

    @property
    def version(self):
        return tuple(f.version for f in self.functions)

    def __call__(self, *inputs):
        mean, err = _gp_predict(self.gp_stack(), concatenate_inputs(inputs))
        return mean.cpu().numpy(), err.cpu().numpy()

    def predict_device(self, points, want_var=False):
        return _gp_predict(self.gp_stack(), points, want_var)

    def torch(self, points):
        """Stacked (mean, beta * sigma), [n, D] each, differentiable (``functions.py:278-291``)."""
        parts = [f.torch(points) for f in self.functions]
        return torch.cat([p[0] for p in parts], dim=1), torch.cat([p[1] for p in parts], dim=1)

    def add_data_point(self, x, y):
        for fun, yi in zip(self.functions, np.asarray(y).squeeze()):
            fun.add_data_point(x, yi)
