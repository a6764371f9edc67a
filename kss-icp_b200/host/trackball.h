// trackball.h -- empty stand-in (SGI trackball of the viewer), unused on the path.
#pragma once
