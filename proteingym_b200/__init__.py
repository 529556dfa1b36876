"""pgscore-b200: B200-native scorer for ProteinGym's PLM log-likelihood hot path (ESM masked-marginals, Tranception, TranceptEVE).
Entry points: compute_fitness / score_tranception_proteingym / score_trancepteve (CLI drop-ins), esm_engine.EsmScorer,
tranception_engine.TranceptionScorer, trancepteve_engine.TranceptEVEScorer, run_assays (multi-GPU driver); C ABI in include/pgscore.h."""
