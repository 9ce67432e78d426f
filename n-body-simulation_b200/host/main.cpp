// Command-line driver with the reference's flags (reference src/main.cpp:9-240, README.md:76-123):
//   mandatory  --file --dt --t_end --vs --vs_dir --algorithm=<naive|BarnesHut>
//   optional   --use_gpus --energy --block_size --opt_stage --theta --num_wi_octree --num_wi_top_octree --num_wi_AABB
//              --num_wi_com --max_level_top_octree --wg_size_barnes_hut --sort_bodies --storage_size_param
//              --stack_size_param
//   new        --stream_output (write each snapshot immediately, O(N) host memory), --vtp_format=<ascii|binary>,
//              --checkpoint=<path> [--checkpoint_every_vs] (binary state for resuming; --file accepts such a state)
// cxxopts (a network FetchContent dependency of the reference) is replaced by the small parser below, which accepts
// `--key=value`, `--key value` and bare boolean flags.  One process drives one GPU; under torchrun-style launchers
// (WORLD_SIZE / RANK / LOCAL_RANK) the processes share the bodies and rank 0 writes the output.
#include <cstdlib>
#include <iostream>
#include <map>
#include <set>
#include <stdexcept>
#include <string>

#include "BarnesHutAlgorithm.hpp"
#include "Configuration.hpp"
#include "InputParser.hpp"
#include "NaiveAlgorithm.hpp"
#include "SimulationData.hpp"
#include "TimeConverter.hpp"

namespace {

class Options {
public:
    Options(int argc, char **argv) {
        static const std::set<std::string> known = {
            "file", "dt", "t_end", "vs", "vs_dir", "theta", "num_wi_octree", "num_wi_top_octree", "num_wi_AABB",
            "num_wi_com", "max_level_top_octree", "storage_size_param", "stack_size_param", "block_size", "algorithm",
            "energy", "sort_bodies", "use_gpus", "wg_size_barnes_hut", "opt_stage", "stream_output", "vtp_format",
            "checkpoint", "checkpoint_every_vs"};
        static const std::set<std::string> booleans = {"energy", "sort_bodies", "use_gpus", "stream_output",
                                                       "checkpoint_every_vs"};
        for (int i = 1; i < argc; ++i) {
            std::string arg = argv[i];
            if (arg.rfind("--", 0) != 0) throw std::invalid_argument("unexpected argument " + arg);
            arg.erase(0, 2);
            std::string key = arg, value;
            bool has_value = false;
            const std::size_t eq = arg.find('=');
            if (eq != std::string::npos) {
                key = arg.substr(0, eq);
                value = arg.substr(eq + 1);
                has_value = true;
            }
            if (!known.count(key)) throw std::invalid_argument("Option '" + key + "' does not exist");
            if (!has_value) {
                if (booleans.count(key) && (i + 1 >= argc || std::string(argv[i + 1]).rfind("--", 0) == 0)) {
                    value = "true";
                } else if (i + 1 < argc) {
                    value = argv[++i];
                } else {
                    throw std::invalid_argument("Option '" + key + "' is missing an argument");
                }
            }
            values[key] = value;
        }
    }
    std::size_t count(const std::string &k) const { return values.count(k); }
    const std::string &str(const std::string &k) const {
        auto it = values.find(k);
        if (it == values.end()) throw std::invalid_argument("Option '" + k + "' not present");
        return it->second;
    }
    int integer(const std::string &k) const {
        std::size_t used = 0;
        const std::string &s = str(k);
        int v;
        try { v = std::stoi(s, &used); } catch (const std::exception &) { used = 0; v = 0; }
        if (used != s.size() || s.empty()) throw std::invalid_argument("Argument '" + s + "' failed to parse for option '" + k + "'");
        return v;
    }
    double real(const std::string &k) const {
        std::size_t used = 0;
        const std::string &s = str(k);
        double v;
        try { v = std::stod(s, &used); } catch (const std::exception &) { used = 0; v = 0; }
        if (used != s.size() || s.empty()) throw std::invalid_argument("Argument '" + s + "' failed to parse for option '" + k + "'");
        return v;
    }
    bool boolean(const std::string &k) const {
        const std::string &s = str(k);
        if (s == "true" || s == "1" || s == "t" || s == "True") return true;
        if (s == "false" || s == "0" || s == "f" || s == "False") return false;
        throw std::invalid_argument("Argument '" + s + "' failed to parse for option '" + k + "'");
    }

private:
    std::map<std::string, std::string> values;
};

int envInt(const char *name, int fallback) {
    const char *v = std::getenv(name);
    return v ? std::atoi(v) : fallback;
}

}  // namespace

int main(int argc, char *argv[]) {
    try {
        Options options(argc, argv);

        std::string path = options.str("file");
        std::string outputDirectoryPath = options.str("vs_dir");
        std::string dt_input = options.str("dt"), t_end_input = options.str("t_end"), vs_input = options.str("vs");
        const double dt = TimeConverter::convertToEarthDays(dt_input);
        const double t_end = TimeConverter::convertToEarthDays(t_end_input);
        const double visualizationStepWidth = TimeConverter::convertToEarthDays(vs_input);
        const std::string algorithm = options.str("algorithm");

        SimulationData simulationData;
        InputParser::parse_input(path, simulationData);
        if (simulationData.mass.empty()) throw std::invalid_argument("input file contains no bodies: " + path);

        const int storageSizeParam = options.count("storage_size_param") ? options.integer("storage_size_param") : 16;
        const int stackSizeParam = options.count("stack_size_param") ? options.integer("stack_size_param") : 16;
        configuration::initializeConfigValues((d_type::int_t) simulationData.mass.size(), storageSizeParam, stackSizeParam);
        configuration::worldSize = envInt("WORLD_SIZE", 1);
        configuration::rank = envInt("RANK", 0);
        configuration::localRank = envInt("LOCAL_RANK", 0);
        const bool talk = configuration::rank == 0;

        if (options.count("energy")) configuration::setEnergyComputation(options.boolean("energy"));
        if (options.count("use_gpus")) configuration::setDeviceGPU(options.boolean("use_gpus"));

        namespace bh = configuration::barnes_hut_algorithm;
        if (algorithm == "naive") {
            if (options.count("block_size")) configuration::setBlockSize(options.integer("block_size"));
            if (options.count("opt_stage")) {
                const int stage = options.integer("opt_stage");
                if (stage > 2 || stage < 0) throw std::invalid_argument("Optimization stage must be 0,1 or 2");
                configuration::setOptimizationStage(stage);
            }
            if (talk) {
                std::cout << "Naive algorithm configuration:" << std::endl;
                std::cout << "Block Size ------------------------------------ " << configuration::naive_algorithm::blockSize << std::endl;
                std::cout << "Optimization stage acceleration kernel -------- " << configuration::naive_algorithm::optimization_stage << std::endl;
                std::cout << std::endl << std::endl;
            }
        } else if (algorithm == "BarnesHut") {
            if (options.count("theta")) configuration::setTheta(options.real("theta"));
            if (options.count("num_wi_octree")) configuration::setOctreeWorkItemCount(options.integer("num_wi_octree"));
            if (options.count("num_wi_top_octree")) configuration::setOctreeTopWorkItemCount(options.integer("num_wi_top_octree"));
            if (options.count("num_wi_com")) configuration::setCenterOfMassWorkItemCount(options.integer("num_wi_com"));
            if (options.count("max_level_top_octree")) configuration::setMaxBuildLevel(options.integer("max_level_top_octree"));
            if (options.count("num_wi_AABB")) configuration::setAABBWorkItemCount(options.integer("num_wi_AABB"));
            if (options.count("sort_bodies")) configuration::setSortBodies(options.boolean("sort_bodies"));
            if (options.count("wg_size_barnes_hut")) configuration::setWorkGroupSizeBarnesHut(options.integer("wg_size_barnes_hut"));
            if (talk) {
                std::cout << "Barnes-Hut algorithm configuration:" << std::endl;
                std::cout << "Theta ----------------------------------- " << bh::theta << std::endl;
                std::cout << "Work-items AABB creation ---------------- " << bh::AABBWorkItemCount << std::endl;
                std::cout << "Work-items octree creation -------------- " << bh::octreeWorkItemCount << std::endl;
                std::cout << "Work-items center of mass calculation --- " << bh::centerOfMassWorkItemCount << std::endl;
                std::cout << "Work-items top of octree creation ------- " << bh::octreeTopWorkItemCount << std::endl;
                std::cout << "Maximum build level top of octree ------- " << bh::maxBuildLevel << std::endl;
                std::cout << "Work-group size acceleration kernel ----- " << bh::workGroupSize << std::endl;
                std::cout << "Sort Bodies enabled --------------------- " << bh::sortBodies << std::endl;
                std::cout << std::endl << std::endl;
            }
        } else {
            throw std::invalid_argument("Algorithm must either be <naive> or <BarnesHut>");
        }

        // new, optional: write every snapshot as soon as it is complete instead of keeping all of them in RAM
        const bool stream = options.count("stream_output") && options.boolean("stream_output");
        bool binaryVtp = false;
        if (options.count("vtp_format")) {
            const std::string &format = options.str("vtp_format");
            if (format != "ascii" && format != "binary") throw std::invalid_argument("vtp_format must either be <ascii> or <binary>");
            binaryVtp = format == "binary";
        }
        const std::string checkpoint = options.count("checkpoint") ? options.str("checkpoint") : std::string();
        const bool checkpointEveryVs = options.count("checkpoint_every_vs") && options.boolean("checkpoint_every_vs");
        auto simulate = [&](nBodyAlgorithm &run) {
            if (stream) run.enableStreaming(simulationData);
            run.setBinaryOutput(binaryVtp);
            run.setCheckpoint(checkpoint, checkpointEveryVs);
            run.startSimulation(simulationData);
            run.generateParaViewOutput(simulationData);
        };
        if (algorithm == "naive") {
            NaiveAlgorithm run(dt, t_end, visualizationStepWidth, outputDirectoryPath);
            simulate(run);
        } else {
            BarnesHutAlgorithm run(dt, t_end, visualizationStepWidth, outputDirectoryPath);
            simulate(run);
        }
    } catch (const std::exception &e) {
        std::cerr << "terminate called after throwing: " << e.what() << std::endl;
        return 1;
    }
    return 0;
}
