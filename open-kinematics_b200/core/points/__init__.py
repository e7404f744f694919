"""Point bookkeeping of the host mirror."""
