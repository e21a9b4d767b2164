"""Helpers the reference keeps under utils/ (/root/reference/utils/utils.py), over the CUDA library."""
