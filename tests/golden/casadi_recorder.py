"""TEST INFRASTRUCTURE (build container only): a RECORDING stand-in for the `casadi` module.

The reference states its optimisation problems through CasADi's Opti stack (control/control.py:203-242, 270-449,
492-597, 640-701; planning/overtake_traj_planner.py:263-364).  CasADi is not installable here, but the *statement* of
a problem does not need a solver: this module implements just enough of the API the reference touches
(`Opti().variable / subject_to / minimize / solver / set_initial / solve / debug.value`, `mtimes`, slicing, `.T`,
arithmetic, comparisons) to RECORD the cost and every constraint as numeric closures over the decision variables.
`solve()` raises RuntimeError -- the reference catches it (or, for mpc_lti, lets it escape) -- and the recorded
problem stays in `LAST`.  tests/golden/make_nlp_golden.py then evaluates the reference's own cost / constraints
at random points and stores the numbers; tests/test_reference_statement.py checks that the problem data our shims
pack (and the oracle's problem functions) give the same numbers.  This pins the PROBLEM STATEMENT to the reference's
code; the solver algorithm (IPOPT) stays unpinned.
"""
import numpy as np

LAST = []          # every Opti() created since the last clear()


def clear():
    del LAST[:]


def _as2d(v):
    a = np.asarray(v, dtype=float)
    if a.ndim == 0:
        return a.reshape(1, 1)
    if a.ndim == 1:
        return a.reshape(-1, 1)          # CasADi treats a flat numpy vector as a column
    return a


class Expr:
    """A numeric closure ctx -> 2-D array.  ctx maps Variable objects to their (n, m) values."""
    __array_ufunc__ = None               # numpy scalars / arrays on the left defer to our reflected operators
    __array_priority__ = 1000

    def __init__(self, fn, shape=None):
        self.fn = fn
        self._shape = shape

    def val(self, ctx):
        return _as2d(self.fn(ctx))

    @property
    def shape(self):
        return self._shape

    @property
    def T(self):
        return Expr(lambda c, s=self: s.val(c).T, None if self._shape is None else self._shape[::-1])

    def __hash__(self):
        return id(self)

    # ---- arithmetic
    @staticmethod
    def _ev(o, c):
        return o.val(c) if isinstance(o, Expr) else _as2d(o)

    def _bin(self, o, f):
        return Expr(lambda c, a=self, b=o: f(Expr._ev(a, c), Expr._ev(b, c)))

    def _rbin(self, o, f):
        return Expr(lambda c, a=o, b=self: f(Expr._ev(a, c), Expr._ev(b, c)))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._rbin(o, np.add)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._rbin(o, np.subtract)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._rbin(o, np.multiply)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __rtruediv__(self, o): return self._rbin(o, np.divide)
    def __pow__(self, p): return Expr(lambda c, a=self, p=p: a.val(c) ** p)
    def __neg__(self): return Expr(lambda c, a=self: -a.val(c))

    # ---- comparisons record constraints
    def __eq__(self, o): return Con(self, "==", o)
    def __le__(self, o): return Con(self, "<=", o)
    def __ge__(self, o): return Con(self, ">=", o)

    # ---- indexing: the result stays 2-D, as in CasADi
    def __getitem__(self, idx):
        if not isinstance(idx, tuple):
            idx = (idx, slice(None))

        def norm(i):
            if isinstance(i, slice):
                return i
            i = int(i)
            return slice(i, i + 1 if i != -1 else None)
        r, cc = norm(idx[0]), norm(idx[1])
        return Expr(lambda c, a=self, r=r, cc=cc: a.val(c)[r, cc])


class Variable(Expr):
    def __init__(self, n, m):
        super().__init__(lambda c, s=None: None, (n, m))
        self.fn = lambda c, s=self: c[s]


class Con:
    """lhs (op) rhs with Expr or numeric sides.  residual(ctx): '==' -> lhs - rhs, '>=' / '<=' -> the quantity that must be >= 0."""

    def __init__(self, lhs, op, rhs):
        self.lhs, self.op, self.rhs = lhs, op, rhs

    def residual(self, ctx):
        a, b = Expr._ev(self.lhs, ctx), Expr._ev(self.rhs, ctx)
        d = (a - b) if self.op in ("==", ">=") else (b - a)
        return np.asarray(d, float).ravel(order="F")

    def __bool__(self):
        raise TypeError("a recorded constraint has no truth value")


class _Debug:
    def __init__(self, opti):
        self.opti = opti

    def value(self, v):
        return np.zeros(v.shape) if isinstance(v, Variable) else 0.0


class Opti:
    def __init__(self):
        self.variables, self.constraints, self.cost, self.initial = [], [], None, []
        self.debug = _Debug(self)
        LAST.append(self)

    def variable(self, n=1, m=1):
        v = Variable(int(n), int(m))
        self.variables.append(v)
        return v

    def subject_to(self, con):
        assert isinstance(con, Con), type(con)
        self.constraints.append(con)

    def minimize(self, cost):
        self.cost = cost

    def set_initial(self, var, value):
        self.initial.append((var, value))

    def solver(self, *a, **k):
        pass

    def solve(self):
        raise RuntimeError("casadi_recorder: problems are recorded, not solved")

    # ---- evaluation helpers for the golden generator
    def eval_cost(self, ctx):
        return float(np.asarray(Expr._ev(self.cost, ctx)).sum())

    def eval_constraints(self, ctx):
        eq = [c.residual(ctx) for c in self.constraints if c.op == "=="]
        ine = [c.residual(ctx) for c in self.constraints if c.op != "=="]
        return (np.concatenate(eq) if eq else np.zeros(0)), (np.concatenate(ine) if ine else np.zeros(0))


def mtimes(a, b):
    return Expr(lambda c, a=a, b=b: Expr._ev(a, c) @ Expr._ev(b, c))


__all__ = ["Opti", "mtimes"]
