"""1-GPU check of the functor header after the in-kernel-exchange change: the header's device-resident search (single
GPU: SearchExchange.G = 1) == the host-driven batched search, bit for bit, for both user objectives (prebuilt library)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import fortran_library_b200 as fl
ulib = C.CDLL(os.path.join(ROOT, "tests", "link", "libuser_objective_prebuilt.so"))
ok = True
for which, n in ((0, 1001), (1, 777), (1, 1 << 16)):
    prob = fl.capi.Problem()
    ulib.user_problem(which, C.byref(prob))
    x0 = np.random.default_rng(n + which).uniform(-0.5, 1.5, n)
    res = []
    for dev in (False, True):
        x = x0.copy()
        ob = fl.Observer()
        st = fl.LBFGS(prob, x, Memory=6, Warning=False, MaxIteration=80, observer=ob, device_search=dev)
        res.append((x, ob.rows, st.n_trials, st.n_f_fd, st.n_f, st.n_fd, st.host_syncs, st.n_batched_passes))
    same = np.array_equal(res[0][0], res[1][0]) and res[0][1] == res[1][1] and res[0][2:6] == res[1][2:6]
    print(f"functor {which} n={n}: device-resident == host-driven: {same}; round trips {res[1][6]} vs {res[0][6]}; "
          f"batched passes (host-driven) {res[0][7]}, trials {res[0][2]}")
    ok = ok and same and res[1][6] < res[0][6]
print("OK" if ok else "FAIL")
sys.exit(0 if ok else 1)
