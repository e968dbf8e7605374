// SOIL2/SOIL2.h -- STAND-IN: P3/main.cpp includes it but calls nothing from it.  *** TEST INFRASTRUCTURE, NOT PRODUCT.
