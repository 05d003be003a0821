"""The closed-loop scenarios of tests/test_reference_scenarios_host.py (the reference's
``tests/test_solve_ik.py``) on the GPU, through the unbatched drop-in API and the C-ABI."""

import pytest

from tests import test_reference_scenarios_host as s

pytestmark = pytest.mark.gpu


def test_checks_and_ignores_configuration_limits():
    s.test_checks_and_ignores_configuration_limits()


def test_no_task_gives_zero_velocity():
    s.test_no_task_gives_zero_velocity()


def test_single_task_fulfilled():
    s.test_single_task_fulfilled()


def test_single_task_convergence():
    s.test_single_task_convergence()


def test_single_task_translation():
    s.test_single_task_translation()


def test_three_tasks_fulfilled_and_convergence():
    s.test_three_tasks_fulfilled()
    s.test_three_tasks_convergence()


def test_com_task_fulfilled_and_convergence():
    s.test_com_task_fulfilled_and_convergence()


def test_model_with_no_joint_limit_has_no_inequalities():
    s.test_model_with_no_joint_limit_has_no_inequalities()


def test_non_finite_target_raises_no_solution_found():
    s.test_non_finite_target_raises_no_solution_found()
