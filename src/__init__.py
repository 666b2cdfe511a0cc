# Package src -- drop-in shims for the reference tree.
