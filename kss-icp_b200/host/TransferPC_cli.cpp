// TransferPC_cli.cpp -- makes a registration test pair the way the reference made data/registration/* (transferPC.hpp):
//   TransferPC_cli <model.xyz> <cord 1|2|3> <angle> [scale=1] [translate=0] [wlopNumber=8000] [gridCell=extent/64]
// writes <model>.wlop (unrotated resample) and <model>.gird (second resample, rotated / scaled / translated).
#include <cstdlib>
#include "transferPC.hpp"

int main(int argc, char** argv) {
	if (argc < 4) { std::cerr << "usage: " << argv[0] << " <model.xyz> <cord 1|2|3> <angle> [scale] [translate] [wlopNumber] [gridCell]\n"; return 2; }
	TransferPC t;
	t.TransferPC_init(argv[1], argc > 6 ? std::atoi(argv[6]) : 8000, argc > 7 ? std::atof(argv[7]) : 0.0);
	t.TransferPC_Transfer(std::atoi(argv[2]), std::atof(argv[3]));
	if (argc > 4 && std::atof(argv[4]) != 1.0) t.TransferPC_Scale(std::atof(argv[4]));
	if (argc > 5 && std::atof(argv[5]) != 0.0) t.TransferPC_Translate(std::atof(argv[5]));
	std::vector<std::vector<std::vector<double>>> r = t.TransferPC_ReturnPoints();
	std::cout << "wrote " << t.filewlop << " (" << r[0].size() << " points) and " << t.filegird << " (" << r[1].size() << " points)" << std::endl;
	return 0;
}
