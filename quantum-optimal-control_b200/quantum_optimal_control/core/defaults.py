"""Convergence defaults of the reference's ``Convergence`` object (core/convergence.py:16-49)."""

CONVERGENCE_DEFAULTS = {
    'rate': 0.01,
    'update_step': 100,
    'evol_save_step': 100,
    'conv_target': 1e-8,
    'max_iterations': 5000,
    'learning_rate_decay': 2500,
    'min_grad': 1e-25,
}


class Convergence:
    """Holds the convergence settings (attribute names as in core/convergence.py:16-58) and the
    recorded history (the reference's matplotlib summary, core/convergence.py:121-222, is out of scope)."""

    def __init__(self, sys_para, time_unit, convergence):
        self.sys_para = sys_para
        self.time_unit = time_unit
        for key, default in CONVERGENCE_DEFAULTS.items():
            setattr(self, key, convergence[key] if key in convergence else default)
        self.reset_convergence()

    def reset_convergence(self):
        self.costs, self.reg_costs, self.iterations, self.learning_rate = [], [], [], []
        self.last_iter = 0
        self.accumulate_rate = 1.00

    def record(self, cost, reg_cost):
        self.costs.append(cost)
        self.reg_costs.append(reg_cost)
        self.iterations.append(self.last_iter)
        self.last_iter += self.update_step
