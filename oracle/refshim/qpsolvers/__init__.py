"""Stand-in for the ``qpsolvers`` module NAME (see ../README.md).  TEST INFRASTRUCTURE.
``solve_problem`` runs the oracle's fp64 Goldfarb-Idnani solver (``oracle/qp.py``), i.e. the
algorithm ``solver="quadprog"`` names, whatever ``solver=`` says."""
import numpy as np

from oracle import qp as _qp

available_solvers = ["quadprog"]


class Problem:
    def __init__(self, P, q, G=None, h=None, A=None, b=None, lb=None, ub=None):
        self.P, self.q, self.G, self.h, self.A, self.b, self.lb, self.ub = P, q, G, h, A, b, lb, ub

    def unpack(self):
        return self.P, self.q, self.G, self.h, self.A, self.b, self.lb, self.ub


class Solution:
    def __init__(self, problem):
        self.problem = problem
        self.x = None
        self.z = None
        self.y = None
        self.found = False
        self.extras = {}


def solve_problem(problem, solver=None, **kwargs):
    assert problem.lb is None and problem.ub is None
    res = _qp.solve_qp(problem.P, problem.q, problem.G, problem.h, problem.A, problem.b)
    sol = Solution(problem)
    sol.found = bool(res.found)
    sol.x = None if not res.found else np.array(res.x)
    sol.extras = {"iterations": getattr(res, "iterations", None)}
    return sol


def solve_qp(P, q, G=None, h=None, A=None, b=None, lb=None, ub=None, solver=None, **kwargs):
    return solve_problem(Problem(P, q, G, h, A, b, lb, ub), solver=solver).x
