#!/bin/bash
# Runs every GPU test function in its own process (a faulting kernel must not poison the rest)
# and writes one log per function under gpurun_out/.  Usage: tests/run_gpu_suite.sh [kernels|step|all]
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
what=${1:-all}
run() { # file, function
  timeout 600 python -m pytest "tests/$1" -m gpu -q -k "$2" -p no:cacheprovider 2>&1 | tail -40 > "gpurun_out/t_$2.log"
  echo "== $2: $(tail -1 gpurun_out/t_$2.log)"
}
if [ "$what" = kernels ] || [ "$what" = all ]; then
  for f in test_conv_tensor_core_path test_conv_forward_dgrad_wgrad test_conv_rejects_bad_descriptor test_batchnorm_helpers test_layout_transposes \
           test_linear_forward_backward test_latent_sample_kl test_gain_stage test_gain_reports_non_pd \
           test_fused_recon_loss test_fused_adam_matches_torch test_gp_posterior_matches_oracle; do
    run test_gpu_kernels.py $f
  done
fi
if [ "$what" = step ] || [ "$what" = all ]; then
  for f in test_step_matches_oracle_and_golden test_step_tensor_core_mode_within_bf16_tolerance test_drop_in_training_loop_decreases_loss test_properties_at_baseline_batch \
           test_ragged_last_batch_and_single_volume test_forward_without_injected_noise_uses_reference_rng_order test_no_cpu_fallback; do
    run test_gpu_step.py $f
  done
fi
