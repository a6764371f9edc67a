// Method_Octree.hpp -- empty stand-in: included by Main_KSS_ICP.cpp:31, never called on the path.
#pragma once
