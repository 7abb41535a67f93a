// mcac_b200 — `MCAC <params.ini>`: same command line and exit codes as the reference's src/main.cpp:26-56.
#include <filesystem>
#include <fstream>
#include <iostream>

#include "aggregat_list.hpp"
#include "physical_model.hpp"

int main(int argc, char *argv[]) {
    if (argc <= 1) {
        std::cout << "Missing argument : param file." << std::endl;
        return mcac::INPUT_ERROR;
    }
    try {
        mcac::PhysicalModel physicalmodel(argv[1]);
        // output_dir is relative to the directory of the .ini (physical_model.cpp:195); created if missing
        namespace fs = std::filesystem;
        const fs::path out = fs::absolute(fs::path(argv[1])).parent_path() / physicalmodel.output_dir;
        fs::create_directories(out);
        physicalmodel.output_dir = out.string();
        {   // the echo of the parsed parameters (physical_model.cpp:271-272)
            std::ofstream os(out / "params.ini");
            os << physicalmodel.ini_echo;
        }
        mcac::AggregatList aggregates(&physicalmodel);
        mcac::calcul(physicalmodel, aggregates);
    } catch (const mcac::BaseException &e) {
        std::cerr << e.what() << std::endl;
        return e.code;
    } catch (const std::exception &e) {
        std::cerr << e.what() << std::endl;
        return mcac::UNKNOWN_ERROR;
    }
    return mcac::NO_ERROR;
}
