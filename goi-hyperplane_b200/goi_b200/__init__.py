"""Host-side helpers around the drop-in rasterizer: synthetic scenes (bench/test inputs), the
semantic-hyperplane mask front end, and view-sharded data parallelism."""
