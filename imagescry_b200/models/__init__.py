"""Models of the sift path (mirror of `imagescry.models`)."""
