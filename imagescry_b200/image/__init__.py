"""Image transforms of the sift path (mirror of `imagescry.image`)."""
