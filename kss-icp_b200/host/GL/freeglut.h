// GL/freeglut.h -- empty stand-in so Main_KSS_ICP.cpp:22 compiles without the viewer dependency.
#pragma once
