"""Stand-in for scikit-image (absent from this image): only the names the reference imports at module level.
None of them is on the model's forward path; the PSNR here restates utils.py:644-659 / skimage's definition."""
