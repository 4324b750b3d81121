"""Mirrors of the reference's config_conditional modules (Conditional BRUNO: m1_shapenet and its drivers)."""
