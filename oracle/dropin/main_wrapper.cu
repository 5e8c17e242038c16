// Compiles the reference's main.cu (or example/main.cu) unchanged against this repo's header;
// see test_wrapper.cu for the include-guard mechanism.
#include <tensor.cuh>
#include REF_SOURCE_FILE
