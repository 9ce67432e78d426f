#include "TimeMeasurement.hpp"

#include <fstream>

void TimeMeasurement::addTimingSequence(const std::string &name) { times[name] = {}; }

void TimeMeasurement::addTimeToSequence(const std::string &sequenceName, double time) {
    times[sequenceName].push_back(time);
}

void TimeMeasurement::setProperties(std::string &algorithm, d_type::int_t &bodyCountArg, std::string &deviceArg) {
    algorithmType = algorithm;
    bodyCount = bodyCountArg;
    device = deviceArg;
}

void TimeMeasurement::exportJSON(const std::string &path) {
    namespace bh = configuration::barnes_hut_algorithm;
    std::ofstream js(path);
    auto field = [&js](const char *key) -> std::ofstream & {
        js << "  \"" << key << "\": ";
        return js;
    };
    js << "{ \n";
    field("algorithm") << "\"" << algorithmType << "\",\n";
    field("device") << "\"" << device << "\",\n";
    if (algorithmType == "Naive Algorithm") {
        field("block size") << configuration::naive_algorithm::blockSize << ",\n";
        field("optimization stage") << configuration::naive_algorithm::optimization_stage << ",\n";
    } else {
        field("theta") << bh::theta << ",\n";
        field("work-group size acceleration") << bh::workGroupSize << ",\n";
        field("work-items AABB") << bh::AABBWorkItemCount << ",\n";
        field("work-items octree") << bh::octreeWorkItemCount << ",\n";
        field("work-items center of mass") << bh::centerOfMassWorkItemCount << ",\n";
        field("work-items top octree") << bh::octreeTopWorkItemCount << ",\n";
        field("max build-level top octree") << bh::maxBuildLevel << ",\n";
        field("bodies sorted") << bh::sortBodies << ",\n";
    }
    field("body count") << bodyCount;
    for (const auto &entry : times) {
        const std::vector<double> &seq = entry.second;
        if (seq.empty()) continue;
        js << ",\n  \"" << entry.first << "\": [";
        for (std::size_t t = 0; t < seq.size(); ++t) js << (t ? ", " : "") << seq[t];
        js << "]";
    }
    js << "\n}";
}
