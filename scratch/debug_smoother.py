"""fixed-lag smoother on the 1-d linear-Gaussian model: error of the trajectory means against the RTS smoother per time index"""
import sys, numpy as np
sys.path.insert(0, '/root/repo')
import mocat_b200 as mocat
F, Q, R, P0 = 0.9, 0.5, 0.8, 1.0
sc = mocat.ssm.TimeHomogenousLinearGaussian(np.zeros(1), [[P0]], [[F]], [[Q]], [[1.0]], [[R]])
T, lag = 22, 6
sim = sc.simulate(np.arange(float(T)), 4)
mu, P, mus, Ps, mup, Pp = 0.0, P0, [], [], [], []
for k in range(T):
    if k > 0:
        mu, P = F * mu, F * P * F + Q
    mup.append(mu); Pp.append(P)
    K = P / (P + R)
    mu, P = mu + K * (sim.y[k, 0] - mu), (1 - K) * P
    mus.append(mu); Ps.append(P)
sm = mus[:]
for k in range(T - 2, -1, -1):
    G = Ps[k] * F / Pp[k + 1]
    sm[k] = mus[k] + G * (sm[k + 1] - mup[k + 1])
sm = np.array(sm)
pf = mocat.ssm.BootstrapFilter()
for bs in (True, False):
    for n, key in ((3000, 9), (3000, 10), (1500, 11), (6000, 12)):
        p = mocat.ssm.initiate_particles(sc, pf, n, key, y=sim.y[0], t=sim.t[0])
        for k in range(1, T):
            p = mocat.ssm.propagate_particle_smoother(sc, pf, p, sim.y[k], sim.t[k], 100 * key + k, lag, backward_sim=bs)
        err = p.value[:, :, 0].mean(1) - sm
        print("bs" if bs else "pf", n, key, "max %.3f at %d" % (np.abs(err).max(), np.abs(err).argmax()), np.round(err, 2),
              "uniq", len(np.unique(p.value[T // 2, :, 0])), flush=True)
    if bs:
        mf = p.marginal_filter
        w = np.exp(mf.log_weight - mf.log_weight.max(1, keepdims=True)); w /= w.sum(1, keepdims=True)
        print("filter mean err", np.round((w * mf.value[:, :, 0]).sum(1) - np.array(mus), 2))
        full = mocat.ssm.backward_simulation(sc, mf, 77, n)
        print("full FFBSi err", np.round(full.value[:, :, 0].mean(1) - sm, 2))
