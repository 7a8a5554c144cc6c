"""Empty stub (TEST INFRASTRUCTURE ONLY): geopandas is imported but unused on the smooth_fit path."""
