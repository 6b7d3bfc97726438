"""Drop-ins for the reference's modelling/ package (backbones, Baseline, CTL model step)."""
