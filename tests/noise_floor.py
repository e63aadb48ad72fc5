"""The reference's own rounding noise, per half-band design: RMS distance of its FP64 DF-II sequence from the exact
value of the same recurrence (binary128, oracle.hilbert_truth), relative to the RMS of the analytic signal; white-noise
input.  Identical for Kahan and baseline summation to two digits.  tests/test_truth_pin.py re-measures it on CPU;
scan mode sits within 3e-15 of the truth, so this is also its distance from the reference."""
REF_NOISE_RMS = {0: 1.4e-9, 1: 5.3e-5, 2: 2.7e-5, 3: 1.9e-4, 4: 1.8e-3, 5: 1.5e-3}
