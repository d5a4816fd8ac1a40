#!/bin/bash
# compute-sanitizer over the smoke configuration and a small fit that walks the step kernel's warp-private
# ring / Monte-Carlo queue (count slots overwritten in place between two __syncwarp), the frozen-event paths and
# the gathered sub-fits.  Output -> gpurun_out/sanitizer_<tool>.log (summaries are copied to profiles/).
#   bash scripts/sanitize.sh            # memcheck + racecheck (+ synccheck, initcheck)
mkdir -p gpurun_out
cat > /tmp/brie_sanitize_case.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np, torch
import __graft_entry__ as g
g.smoke()
from brie_b200.models import fit_BRIE_matrix
from util import make_problem, make_lrt_problem
# wide covariates (2 events per lane) + gene features + per-cell intercept: every reduction path of the step kernel
data, effLen, Xc, Xg = make_problem(70, 45, 9, 5, False, 2, seed=3)
fit_BRIE_matrix([x.copy() for x in data], Xc=Xc, Xg=Xg, intercept_mode='cell', LRT_index=[], min_iter=12, max_iter=12,
                MC_size=3, n_eval=3)
# the two-rows-per-iteration form (Kc 11 -> 16 without gene features; Kc 8 with gene features) and a wide design (GEMM form)
for kc, kg, mode in ((11, 0, 'gene'), (8, 3, 'gene'), (20, 0, 'gene')):
    data, effLen, Xc, Xg = make_problem(70, 45, kc, kg, True, 3, seed=4)
    fit_BRIE_matrix([x.copy() for x in data], Xc=Xc, Xg=Xg if kg else None, effLen=effLen, intercept_mode=mode, LRT_index=[],
                    min_iter=12, max_iter=12, MC_size=3, n_eval=3)
# batched LRT with convergence groups: extension rounds in place (active-block list) and on gathered sub-fits
from brie_b200.engine import FitEngine
from oracle.brie2_oracle import add_pseudo_count
data, effLen, Xc, _ = make_lrt_problem(40, 64, seed=3)
add_pseudo_count(data, np.float32(0.01))
eng = FitEngine(data, effLen=effLen, Xc=Xc, masks=[[0], []], model_ids=[0, 1], MC_size=2, seed=1, group_size=8, trace_cap=16)
eng.init_params(); eng.begin_stage(0.01); eng.run_steps(6, 0)
act = np.zeros((2, eng.n_groups), bool); act[0, 1] = act[0, 5] = act[1, 2] = True
assert eng.run_steps_gathered(act, 5, force=True)
eng.set_active_groups(act); eng.run_steps(5, 0)
eng.set_active_groups(np.ones_like(act)); eng.run_steps(2, 0)
lg = eng.eval_loss_gene(6); post = eng.posterior(1)
torch.cuda.synchronize()
print("sanitize case done")
PY
for tool in ${TOOLS:-memcheck racecheck synccheck initcheck}; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 \
      python /tmp/brie_sanitize_case.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool exit=$?" | tee -a gpurun_out/sanitizer_$tool.log
  tail -4 gpurun_out/sanitizer_$tool.log
done
