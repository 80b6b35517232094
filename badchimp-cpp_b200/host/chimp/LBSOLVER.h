// Umbrella header, the counterpart of the reference's src/LBSOLVER.h for the hot path.
#ifndef CHIMP_LBSOLVER_LIB
#define CHIMP_LBSOLVER_LIB
#include "LBglobal.h"
#include "LBlattices.h"
#include "LBfield.h"
#include "LBvtk.h"
#include "LBgrid.h"
#include "LBhalfwaybb.h"
#include "LBpressurebnd.h"
#include "LBbndmpi.h"
#include "LBcollision.h"
#include "LBgpu.h"
#include "LBranks.h"
#include "Input.h"
#include "Output.h"
#endif
