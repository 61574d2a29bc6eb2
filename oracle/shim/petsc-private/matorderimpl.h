/* stand-in for <petsc-private/matorderimpl.h>; hslmc64.c needs nothing from it. */
#include <petscsys.h>
