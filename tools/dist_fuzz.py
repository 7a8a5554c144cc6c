"""Differential check of the package's multi-rank HOST logic against the unmodified reference (build container only).
Random cases of tools/diff_fuzz.py are fitted by WORLD gloo ranks (tests/two_rank_worker.py: kernel wrappers replaced by
tests/host_double.py, every rank assembles its block of rows, the statistics are all-reduced, per-point outputs gathered) and
by the single-process reference; every rank must return the reference's model, final edit mask, z_est and count map, and
take the reference's number of edit iterations.
    python tools/dist_fuzz.py [n_cases=8] [seed=0] [world=2] [widened] [more]"""
import os, pickle, subprocess, sys, tempfile, warnings
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tools'))
import diff_fuzz as F
from oracle import ref_harness as H


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    world = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    rng = np.random.default_rng(seed)
    R = H.load()
    worker = os.path.join(ROOT, 'tests', 'two_rank_worker.py')
    bad = n_run = 0
    for k in range(n_cases):
        pts, kw = F.one_case(rng, k)
        if 'widened' in sys.argv[4:]:
            pts, kw = F.widen(rng, pts, kw)
        tag = 'case %d: n=%d %s' % (k, pts['x'].size, {a: (b if a not in ('mask_data', 'constraint_scaling_maps', 'avg_masks', 'prior_args') else sorted(b))
                                               for a, b in kw.items() if a not in ('W', 'ctr')})
        import sparseqr                     # the stub of oracle/ref_stubs: count the reference's solves = its edit iterations
        solves, orig_solve = [], sparseqr.solve

        def count_solve(*a, **k2):
            solves.append(1)
            return orig_solve(*a, **k2)
        sparseqr.solve = count_solve
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            try:
                S_r = F.reference_fit(R, pts, kw)
            except Exception as e:          # noqa: BLE001 -- error behaviour is diff_fuzz.py's subject
                print('skip ' + tag + '  (reference raises %s)' % type(e).__name__)
                continue
            finally:
                sparseqr.solve = orig_solve
        if S_r['data'] is None or kw['max_iterations'] == 0:
            print('skip ' + tag + '  (no valid data / no iterations)')
            continue
        n_run += 1
        with tempfile.TemporaryDirectory() as tmp:
            case = os.path.join(tmp, 'case.pkl')
            with open(case, 'wb') as f:
                pickle.dump((pts, F.package_kw(kw)), f)
            port = str(34000 + (os.getpid() + 7 * k) % 2000)
            outs = [os.path.join(tmp, 'rank%d.npz' % r) for r in range(world)]
            procs = [subprocess.Popen([sys.executable, worker, str(r), str(world), port, case, outs[r]], stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT) for r in range(world)]
            logs = []
            for p in procs:
                try:
                    logs.append(p.communicate(timeout=600)[0].decode(errors='replace'))
                except subprocess.TimeoutExpired:
                    for q in procs:
                        q.kill()
                    logs.append('TIMEOUT')
            if any(p.returncode != 0 for p in procs):
                print('DIFF ' + tag + '  a rank failed:\n' + '\n'.join(l[-1500:] for l in logs))
                bad += 1
                continue
            res = [np.load(o) for o in outs]
        tse = np.packbits(np.asarray(S_r['data'].three_sigma_edit, dtype=bool))
        m_r = S_r['m']['all']
        failed = []
        for r, out in enumerate(res):
            checks = dict(m0=np.linalg.norm(out['m0'] - m_r) <= 1e-8 * np.linalg.norm(m_r), mask=np.array_equal(out['in_TSE'], tse),
                          z_est=np.max(np.abs(out['z_est'] - np.asarray(S_r['data'].z_est))) < 1e-6,
                          count=np.allclose(out['count'], S_r['m']['z0'].count, equal_nan=True),
                          n_outer=int(out['n_solves']) == len(solves))
            failed += ['rank%d.%s' % (r, c) for c, v in checks.items() if not v]
        if not np.array_equal(res[0]['m0'], res[-1]['m0']):
            failed.append('ranks disagree on m0')
        print(('ok   ' if not failed else 'DIFF ') + tag + ('' if not failed else '  FAILED: %s (edit iterations: reference %d, ranks %s)' % (failed, len(solves), [int(o['n_solves']) for o in res])))
        bad += bool(failed)
    print('%d of %d cases differ (%d of %d drawn cases were run)' % (bad, n_run, n_run, n_cases))
    return bad


if __name__ == '__main__':
    sys.exit(1 if main() else 0)
