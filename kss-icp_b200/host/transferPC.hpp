// transferPC.hpp -- empty stand-in: the reference's test-data generator (CGAL), not on the path.
#pragma once
