"""Row containers for the metric columns the device computes."""
