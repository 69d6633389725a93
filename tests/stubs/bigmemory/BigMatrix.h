// tests/stubs/bigmemory/BigMatrix.h -- NOT bigmemory: the four members of BigMatrix the shim touches, for -fsyntax-only.
#pragma once
struct BigMatrix {
    int matrix_type() const { return 8; }
    void *matrix() { return 0; }
    long nrow() const { return 0; }
    long ncol() const { return 0; }
};
