"""Empty stub (TEST INFRASTRUCTURE ONLY): requests is imported but unused on the smooth_fit path."""
