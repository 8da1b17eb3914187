/* Minimal declarations of the MEX C API -- the part matlab/gpz_b200_mex.cpp uses -- so that the gateway can be compiled and,
 * linked with the mock implementation in mex_mock.cpp, EXECUTED in a container without MATLAB (tests/test_mex_gateway.py).
 * Not a MATLAB header. */
#ifndef MEX_STUB_H
#define MEX_STUB_H
#include <stddef.h>
#include <stdint.h>
typedef struct mxArray_tag mxArray;
typedef size_t mwSize;
typedef bool mxLogical;
typedef enum { mxREAL = 0, mxCOMPLEX = 1 } mxComplexity;
typedef enum { mxDOUBLE_CLASS = 6 } mxClassID;
extern "C" {
double* mxGetPr(const mxArray*);
double mxGetScalar(const mxArray*);
size_t mxGetM(const mxArray*);
size_t mxGetN(const mxArray*);
size_t mxGetNumberOfElements(const mxArray*);
bool mxIsEmpty(const mxArray*);
bool mxIsChar(const mxArray*);
bool mxIsStruct(const mxArray*);
bool mxIsLogical(const mxArray*);
mxLogical* mxGetLogicals(const mxArray*);
mxArray* mxGetField(const mxArray*, size_t, const char*);
int mxGetString(const mxArray*, char*, size_t);
mxArray* mxCreateDoubleMatrix(size_t, size_t, mxComplexity);
mxArray* mxCreateDoubleScalar(double);
mxArray* mxCreateNumericArray(size_t, const mwSize*, mxClassID, mxComplexity);
void mxDestroyArray(mxArray*);
void mexErrMsgIdAndTxt(const char*, const char*, ...);
int mexPrintf(const char*, ...);
int mexEvalString(const char*);
void mexLock(void);
int mexAtExit(void (*)(void));
void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]);   /* as MATLAB's mex.h declares it: C linkage */
}
#endif
