// View.h -- empty stand-in: the reference's GLUT callback prototypes (PS_AIS_Simplification/View.h) belong to
// its viewer, which is outside the registration path.
#pragma once
