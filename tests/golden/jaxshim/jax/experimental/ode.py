"""odeint(func, y0, t, *args): SciPy's Dormand-Prince 5(4) (the method of jax.experimental.ode), dense at t."""
import numpy as np
from scipy.integrate import solve_ivp


def odeint(func, y0, t, *args, rtol=1.4e-8, atol=1.4e-8, mxstep=None, hmax=None):
    t = np.asarray(t, np.float64)
    sol = solve_ivp(lambda tt, yy: func(yy, tt, *args), (t[0], t[-1]), np.asarray(y0, np.float64), t_eval=t, method="RK45",
                    rtol=rtol, atol=atol)
    return sol.y.T
