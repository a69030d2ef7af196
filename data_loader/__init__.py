"""Drop-in for the one `data_loader` entry point the inference callers use (test_code/inference.py, app.py):
`video_this_that_dataset.get_thisthat_sam`. The training datasets of the reference are out of scope."""
