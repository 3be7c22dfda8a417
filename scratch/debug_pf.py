import numpy as np, sys
sys.path.insert(0, '/root/repo')
import mocat_b200 as mocat
dim = 5
ssm = mocat.ssm.TimeHomogenousLinearGaussian(initial_mean=np.zeros(dim), initial_covariance=np.eye(dim),
                                             transition_matrix=np.eye(dim), transition_covariance=np.eye(dim),
                                             likelihood_matrix=np.eye(dim), likelihood_covariance=np.eye(dim))
t = np.arange(20, dtype=np.float64)
sim = ssm.simulate(t, random_key=0)
pf = mocat.ssm.run_particle_filter_for_marginals(ssm, mocat.ssm.BootstrapFilter(), sim.y, t, random_key=0, n=2000)
p = mocat.ssm.initiate_particles(ssm, mocat.ssm.BootstrapFilter(), 2000, 0, sim.y[0], t[0])
for i in range(1, 6):
    p = mocat.ssm.propagate_particle_filter(ssm, mocat.ssm.BootstrapFilter(), p, sim.y[i], t[i], 0)
print("ess batch ", pf.ess[:6]); print("ess online", p.ess[:6]); print("resampled batch", pf.resampled[:6])
for i in range(6):
    print(i, np.abs(p.value[i] - pf.value[i]).max(), np.abs(p.log_weight[i] - pf.log_weight[i]).max())
