"""CPU oracle (test infrastructure only) -- see ssq_oracle.py / ssq_stft_ref.c."""
