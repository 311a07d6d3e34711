"""CPU oracle for the CTC best-path hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``kokoro-align_b200/`` imports this package.  See ``ctc_oracle.c`` (per-cell C
restatement) and ``ctc_oracle_np.py`` (per-frame numpy restatement that mirrors the
reference's own numpy structure and therefore its speed).
"""
