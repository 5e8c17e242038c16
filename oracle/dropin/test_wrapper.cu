// Compiles the reference's test/testTensor.cu UNCHANGED against this repo's include/tensor.cuh.
// The test file includes "../include/tensor.cuh" relative to its own directory, which would pick up the
// reference header; both headers use the include guard TENSOR_CUH, so including ours first turns the
// reference header into an empty file. No reference source is copied or edited.
#include <gtest/gtest.h>
#include <tensor.cuh>
#ifndef TENSOR_CUH
#error "include/tensor.cuh must define TENSOR_CUH"
#endif
#include <numeric>
#include REF_SOURCE_FILE
