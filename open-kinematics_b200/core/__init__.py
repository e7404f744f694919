"""Host mirror of the reference's ``kinematics.core`` solve-path modules."""
