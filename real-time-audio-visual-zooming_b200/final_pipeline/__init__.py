"""Mirror of the reference's Final_pipeline package: `config`, `inference.enhance_audio`, `metrics`,
`batch_run.run_batch`, `run.main` keep their signatures (Final_pipeline/run.py:5, batch_run.py:12,
src/inference.py:144, src/metrics.py:102,125)."""
