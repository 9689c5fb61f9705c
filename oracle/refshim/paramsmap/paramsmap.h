// Stand-in for the reference's (absent, unreachable) lib/paramsmap submodule.
// Only the declaration is needed by the translation units the oracle compiles
// (include/driver/driver.h:52 of the reference names the type in a signature).
// TEST INFRASTRUCTURE — used only by oracle/Makefile.
#ifndef QRB200_ORACLE_PARAMSMAP_STANDIN_H
#define QRB200_ORACLE_PARAMSMAP_STANDIN_H
class ParamsMap;
#endif
