#!/bin/bash
# FFN-phase micro experiments on config B: solo vs pair, with loads or MMA disabled (results garbage; timing only)
for pair in 0 1; do for dbg in 0 1 2 3; do
  echo "== PAIR=$pair DBG=$dbg (1=no TMA, 2=no MMA)"
  FM_PAIR=$pair FM_DBG_FLAGS=$dbg FM_FUSED_COMBINE=0 python scripts/diag.py --cfg B --stage time 2>&1 | grep "phase ffn"
done; done
