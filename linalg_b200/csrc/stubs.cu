// (all operations implemented; file kept empty on purpose so that the build list stays stable)
