// dca-exe: the reference's executable (src/Main.hs:10-22) over libdca_b200.
#include <cstdarg>
#include <iostream>

#include "dca_host.hpp"

int main(int argc, char** argv) {
  try {
    std::random_device rd;
    const uint64_t seed = ((uint64_t)rd() << 32) | rd();  // randomIO :: IO Word64
    const dcahost::Opt opt = dcahost::getOpts(argc, argv, seed);
    const int status = dcahost::runSim(opt, [](const std::string& s) { std::cout << s << std::endl; });
    return status == DCA_SUCCESS || status == DCA_RUNNING ? 0 : 2;
  } catch (const dcahost::Error& e) {
    if (e.code == 0) {  // --help
      std::cout << e.what();
      return 0;
    }
    std::cerr << e.what() << std::endl;
    return 1;
  }
}
