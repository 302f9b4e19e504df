// DTALite_b200 executable: same main() semantics as the reference's flash_dta.cpp:140-246 -- run it in a
// directory holding settings.csv, node.csv, link.csv and the demand file(s).
#include "dtalite_b200.h"
int main() { return dtalite_main(); }
