// C driver around mexFunction for the tests (ctypes): builds mxArrays, calls the gateway the way MATLAB would,
// turns mexErrMsgIdAndTxt (a C++ throw in this runtime) into a return code + id/message, and owns the outputs.
#include <string.h>

#include <string>

#include "mex.h"
#include "mex_runtime.hpp"

namespace {
std::string g_err_id, g_err_msg;
}

#define MEXH_API extern "C" __attribute__((visibility("default")))

MEXH_API mxArray* mexh_numeric(int ndim, const size_t* dims, int is_single, int is_complex) {
  return mxCreateNumericArray(ndim, dims, is_single ? mxSINGLE_CLASS : mxDOUBLE_CLASS, is_complex ? mxCOMPLEX : mxREAL);
}
MEXH_API mxArray* mexh_string(const char* s) { return mxCreateString(s); }
MEXH_API mxArray* mexh_struct(void) { return mxCreateStructMatrix(1, 1, 0, nullptr); }
MEXH_API void mexh_set_field(mxArray* st, const char* name, mxArray* v) { mxSetField(st, 0, name, v); }   // st owns v
MEXH_API void* mexh_data(mxArray* a) {
  if (mxIsDouble(a)) return mxIsComplex(a) ? static_cast<void*>(mxGetComplexDoubles(a)) : static_cast<void*>(mxGetDoubles(a));
  return mxIsComplex(a) ? static_cast<void*>(mxGetComplexSingles(a)) : static_cast<void*>(mxGetSingles(a));
}
MEXH_API int mexh_ndim(const mxArray* a) { return static_cast<int>(mxGetNumberOfDimensions(a)); }
MEXH_API size_t mexh_dim(const mxArray* a, int i) { return mxGetDimensions(a)[i]; }
MEXH_API int mexh_is_single(const mxArray* a) { return mxIsSingle(a); }
MEXH_API int mexh_is_complex(const mxArray* a) { return mxIsComplex(a); }
MEXH_API void mexh_destroy(mxArray* a) { mxDestroyArray(a); }

// 0 = returned normally; 1 = the gateway raised an error (id / message below; plhs left NULL, nothing leaks because the
// gateway creates its outputs before it can fail only through the engine, and MATLAB would free them the same way)
MEXH_API int mexh_call(int nlhs, mxArray** plhs, int nrhs, mxArray** prhs) {
  for (int i = 0; i < nlhs; ++i) plhs[i] = nullptr;
  mxArray* out[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  try {
    mexFunction(nlhs, out, nrhs, const_cast<const mxArray**>(prhs));
  } catch (const minimex::MexError& e) {
    g_err_id = e.id;
    g_err_msg = e.msg;
    for (auto* o : out) mxDestroyArray(o);          // MATLAB destroys arrays created before the error
    return 1;
  }
  for (int i = 0; i < 8; ++i) {
    if (i < nlhs || (i == 0 && out[0])) { if (i < nlhs) plhs[i] = out[i]; else mxDestroyArray(out[i]); }
    else mxDestroyArray(out[i]);
  }
  return 0;
}
MEXH_API const char* mexh_error_id(void) { return g_err_id.c_str(); }
MEXH_API const char* mexh_error_msg(void) { return g_err_msg.c_str(); }
MEXH_API void mexh_clear_mex(void) { minimex::run_atexit(); }      // `clear mex`: registered mexAtExit handlers run
MEXH_API int mexh_atexit_count(void) { return static_cast<int>(minimex::n_atexit()); }
