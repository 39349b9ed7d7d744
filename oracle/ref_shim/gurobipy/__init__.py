"""Import-only gurobipy stand-in + a tiny model layer solved by scipy HiGHS.

TEST INFRASTRUCTURE ONLY.  gurobipy is a proprietary third-party dependency of
the reference (imported at module top of lib/optimUtils.py:1 and
lib/Optimization.py:4) that is not installed here.  The subset below is what
AlonsoMora.assignFromRTVGraph (lib/Optimization.py:359-408) uses: Model,
setParam, tuplelist, tupledict, addVars, setObjective, addConstrs, addConstr,
optimize, status, getAttr('x', vars).  The binary program is handed to
scipy.optimize.milp.  Which optimum is returned when several exist follows
HiGHS, not Gurobi (parity unpinned for ILP ties).
"""
import numpy as np


class _GRB(object):
    BINARY = "B"
    CONTINUOUS = "C"
    MINIMIZE = 1
    MAXIMIZE = -1
    OPTIMAL = 2
    INFEASIBLE = 3

    class Param(object):
        OutputFlag = "OutputFlag"
        LogToConsole = "LogToConsole"
        LogFile = "LogFile"

    class Callback(object):
        MIPSOL = 4


GRB = _GRB()


class LinExpr(object):
    def __init__(self, terms=None, const=0.0):
        self.terms = dict(terms or {})
        self.const = const

    def _add(self, other, sign=1.0):
        out = LinExpr(self.terms, self.const)
        if isinstance(other, Var):
            other = LinExpr({other.idx: 1.0})
        if isinstance(other, LinExpr):
            for k, v in other.terms.items():
                out.terms[k] = out.terms.get(k, 0.0) + sign * v
            out.const += sign * other.const
        else:
            out.const += sign * float(other)
        return out

    def __add__(self, other):
        return self._add(other)

    __radd__ = __add__

    def __sub__(self, other):
        return self._add(other, -1.0)

    def __mul__(self, k):
        return LinExpr({i: v * float(k) for i, v in self.terms.items()}, self.const * float(k))

    __rmul__ = __mul__

    def __eq__(self, rhs):
        return ("==", self, float(rhs))

    def __le__(self, rhs):
        return ("<=", self, float(rhs))

    def __ge__(self, rhs):
        return (">=", self, float(rhs))


class Var(object):
    def __init__(self, idx):
        self.idx = idx
        self.x = 0.0

    def _e(self):
        return LinExpr({self.idx: 1.0})

    def __add__(self, other):
        return self._e() + other

    __radd__ = __add__

    def __mul__(self, k):
        return self._e() * k

    __rmul__ = __mul__


class tuplelist(list):
    pass


def _key_match(key, pattern):
    key = key if isinstance(key, tuple) else (key,)
    return all(p == "*" or p == k for k, p in zip(key, pattern))


class tupledict(dict):
    def sum(self, *pattern):
        expr = LinExpr()
        for k, v in self.items():
            if not pattern or _key_match(k, pattern):
                expr = expr + v
        return expr

    def prod(self, coeffs):
        expr = LinExpr()
        for k, v in self.items():
            if k in coeffs:
                expr = expr + v * coeffs[k]
        return expr


class Model(object):
    def __init__(self, name=""):
        self.nvars = 0
        self.vars = []
        self.obj = LinExpr()
        self.sense = GRB.MINIMIZE
        self.cons = []
        self.status = None

    def setParam(self, *a, **k):
        pass

    def addVars(self, keys, vtype=None, name=""):
        out = tupledict()
        for k in keys:
            v = Var(self.nvars)
            self.nvars += 1
            self.vars.append(v)
            out[k] = v
        return out

    def setObjective(self, expr, sense=GRB.MINIMIZE):
        self.obj = expr if isinstance(expr, LinExpr) else LinExpr() + expr
        self.sense = sense

    def addConstr(self, con, name=""):
        self.cons.append(con)

    def addConstrs(self, gen, name=""):
        for con in gen:
            self.cons.append(con)

    def optimize(self):
        from scipy.optimize import milp, LinearConstraint, Bounds
        n = self.nvars
        c = np.zeros(n)
        for i, v in self.obj.terms.items():
            c[i] = v * (1.0 if self.sense == GRB.MINIMIZE else -1.0)
        rows, lo, hi = [], [], []
        for (op, expr, rhs) in self.cons:
            row = np.zeros(n)
            for i, v in expr.terms.items():
                row[i] = v
            r = rhs - expr.const
            rows.append(row)
            lo.append(r if op in ("==", ">=") else -np.inf)
            hi.append(r if op in ("==", "<=") else np.inf)
        kw = {}
        if rows:
            kw["constraints"] = LinearConstraint(np.array(rows), lo, hi)
        res = milp(c, integrality=np.ones(n), bounds=Bounds(0, 1), **kw)
        if res.status == 0:
            self.status = GRB.OPTIMAL
            for v, x in zip(self.vars, res.x):
                v.x = float(round(x))
            self.objVal = float(res.fun)
        else:
            self.status = GRB.INFEASIBLE

    def getAttr(self, attr, vars_):
        assert attr == "x"
        return tupledict({k: v.x for k, v in vars_.items()})
