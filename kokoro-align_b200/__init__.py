"""B200-native CTC best-path aligner: drop-in for kaiidams/Kokoro-Align's
``kokoro_align.align.ctc_best_path`` / ``best_path`` (align.py:43-124)."""
__version__ = "0.1.0"
