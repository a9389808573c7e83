"""Host-side mirror of the reference's ``Parameters`` (src/parameters.jl:1-39) and of the
status table (src/status.jl:2-23).  Only the fields the live SQP-TR path reads are kept;
``external_optimizer`` names the sub-solver slot the device engine plugs into."""
from __future__ import annotations

APPLICATION_RETURN_STATUS = {
    0: "Solve_Succeeded", 1: "Solved_To_Acceptable_Level", 2: "Infeasible_Problem_Detected",
    3: "Search_Direction_Becomes_Too_Small", 4: "Diverging_Iterates", 5: "User_Requested_Stop",
    6: "Feasible_Point_Found", -1: "Maximum_Iterations_Exceeded", -2: "Restoration_Failed",
    -3: "Error_In_Step_Computation", -4: "Maximum_CpuTime_Exceeded", -5: "Optimize_not_called",
    -6: "Method_not_defined", -10: "Not_Enough_Degrees_Of_Freedom", -11: "Invalid_Problem_Definition",
    -12: "Invalid_Option", -13: "Invalid_Number_Detected", -100: "Unrecoverable_Exception",
    -102: "Insufficient_Memory", -199: "Internal_Error",
}


def moi_termination_status(status: int) -> str:
    """MOI_wrapper.jl:1238-1278."""
    name = APPLICATION_RETURN_STATUS[status]
    table = {
        "Solve_Succeeded": "LOCALLY_SOLVED", "Feasible_Point_Found": "LOCALLY_SOLVED",
        "Infeasible_Problem_Detected": "LOCALLY_INFEASIBLE", "Solved_To_Acceptable_Level": "ALMOST_LOCALLY_SOLVED",
        "Search_Direction_Becomes_Too_Small": "NUMERICAL_ERROR", "Diverging_Iterates": "NORM_LIMIT",
        "User_Requested_Stop": "INTERRUPTED", "Maximum_Iterations_Exceeded": "ITERATION_LIMIT",
        "Maximum_CpuTime_Exceeded": "TIME_LIMIT", "Restoration_Failed": "NUMERICAL_ERROR",
        "Error_In_Step_Computation": "NUMERICAL_ERROR", "Invalid_Option": "INVALID_OPTION",
        "Not_Enough_Degrees_Of_Freedom": "INVALID_MODEL", "Invalid_Problem_Definition": "INVALID_MODEL",
        "Invalid_Number_Detected": "INVALID_MODEL", "Unrecoverable_Exception": "OTHER_ERROR",
    }
    return table.get(name, "MEMORY_LIMIT")


class Parameters:
    """parameters.jl:1-30; ``set_parameter``/``get_parameter`` by field name (:32-39)."""

    def __init__(self, **kw):
        self.algorithm = "SQP-TR"
        self.external_optimizer = "sqpqp-b200"
        self.OutputFlag = 0
        self.StatisticsFlag = 0
        self.tol_direction = 1e-8
        self.tol_residual = 1e-8
        self.tol_infeas = 1e-8
        self.max_iter = 3000
        self.init_mu = 1.0
        self.tr_size = 10.0
        self.use_soc = False
        for k, v in kw.items():
            self.set_parameter(k, v)

    def get_parameter(self, name):
        return getattr(self, name)

    def set_parameter(self, name, val):
        if not hasattr(self, name):
            raise KeyError(f"unknown parameter {name}")
        setattr(self, name, val)
