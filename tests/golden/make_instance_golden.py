"""Records the Ising encodings the reference's own problem classes produce (run in the build
container, where /root/reference is importable):  TSP position encoding
(problems/routing.py:193-328) and SimpleScheduler (problems/simple_scheduler.py:67-127).
The committed .npz files pin tools/instances.py, which re-derives the same J, h vectorised for
the full-size BASELINE configs (cfg4: 64 cities, cfg5: 500 tasks x 100 agents) that the
reference's O(N^2)-per-set_coupling encoder cannot build in reasonable time.

    PYTHONPATH=/root/reference python tests/golden/make_instance_golden.py
"""
import os
import warnings

import numpy as np

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))


def dense(m):
    J = m.couplings.to_dense() if m.couplings.is_sparse else m.couplings
    return J.numpy().astype(np.float32), m.external_fields.numpy().astype(np.float32)


def tsp(n_cities, seed):
    from spin_glass_rl.problems.routing import TSPProblem
    np.random.seed(seed)
    p = TSPProblem()
    p.generate_random_instance(n_locations=n_cities, area_size=100.0)
    m = p.encode_to_ising()
    J, h = dense(m)
    xy = np.array([(l.x, l.y) for l in p.locations], np.float64)
    np.savez_compressed(os.path.join(HERE, f"inst_tsp{n_cities}.npz"), xy=xy, J=J, h=h,
                        seed=seed)


def sched(n_tasks, n_agents, seed):
    from spin_glass_rl.problems.simple_scheduler import SimpleScheduler
    np.random.seed(seed)
    p = SimpleScheduler()
    p.generate_random_instance(n_tasks=n_tasks, n_agents=n_agents)
    m = p.encode_to_ising()
    J, h = dense(m)
    dur = np.array([t.duration for t in p.tasks], np.float64)
    due = np.array([t.due_date if t.due_date else 0.0 for t in p.tasks], np.float64)
    rate = np.array([a.cost_rate for a in p.agents], np.float64)
    np.savez_compressed(os.path.join(HERE, f"inst_sched_{n_tasks}x{n_agents}.npz"), duration=dur,
                        due_date=due, cost_rate=rate, J=J, h=h, seed=seed)


if __name__ == "__main__":
    tsp(5, 4004)
    tsp(6, 4005)
    sched(3, 4, 5005)
    sched(6, 5, 5006)
    print("written")
