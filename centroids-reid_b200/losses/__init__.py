"""Drop-ins for the reference's losses/ package (triplet_loss.py, center_loss.py)."""
