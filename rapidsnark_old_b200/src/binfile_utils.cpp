#include "binfile_utils.hpp"
#include <fcntl.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <stdexcept>
#include <system_error>

namespace BinFileUtils {

BinFile::BinFile(std::string fileName, std::string _type, uint32_t maxVersion)
    : addr(nullptr), size(0), pos(0), version(0), readingSection(nullptr) {
    int fd = open(fileName.c_str(), O_RDONLY);
    if (fd == -1) throw std::system_error(errno, std::generic_category(), "open");
    struct stat sb;
    if (fstat(fd, &sb) == -1) { close(fd); throw std::system_error(errno, std::generic_category(), "fstat"); }
    size = (uint64_t)sb.st_size;
    if (size < 12) { close(fd); throw std::invalid_argument("Invalid file type. It should be " + _type + " and it us <truncated>"); }
    addr = mmap(nullptr, size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    close(fd);
    if (addr == MAP_FAILED) { addr = nullptr; throw std::system_error(errno, std::generic_category(), "mmap"); }

    type.assign((const char *)addr, 4);
    pos = 4;
    if (type != _type)
        throw std::invalid_argument("Invalid file type. It should be " + _type + " and it us " + type);
    version = readU32LE();
    if (version > maxVersion)
        throw std::invalid_argument("Invalid version. It should be <=" + std::to_string(maxVersion) + " and it us " + std::to_string(version));
    uint32_t nSections = readU32LE();
    for (uint32_t i = 0; i < nSections; i++) {
        if (pos + 12 > size) throw std::range_error("Invalid section size");
        uint32_t sType = readU32LE();
        uint64_t sSize = readU64LE();
        if (sSize > size - pos) throw std::range_error("Invalid section size");
        sections[(int)sType].push_back(Section{(void *)((uint8_t *)addr + pos), sSize});
        pos += sSize;
    }
    pos = 0;
}

BinFile::~BinFile() {
    if (addr) munmap(addr, size);
}

void BinFile::startReadSection(uint32_t sectionId, uint32_t sectionPos) {
    if (sections.find((int)sectionId) == sections.end())
        throw std::range_error("Section does not exist: " + std::to_string(sectionId));
    if (sectionPos >= sections[(int)sectionId].size())
        throw std::range_error("Section pos too big. There are " + std::to_string(sections[(int)sectionId].size()) +
                               " and it's trying to access section: " + std::to_string(sectionPos));
    if (readingSection != nullptr) throw std::range_error("Already reading a section");
    pos = (uint64_t)((uint8_t *)sections[(int)sectionId][sectionPos].start - (uint8_t *)addr);
    readingSection = &sections[(int)sectionId][sectionPos];
}

void BinFile::endReadSection(bool check) {
    if (check && readingSection) {
        if ((uint64_t)((uint8_t *)readingSection->start - (uint8_t *)addr) + readingSection->size != pos)
            throw std::range_error("Invalid section size");
    }
    readingSection = nullptr;
}

void *BinFile::getSectionData(uint32_t sectionId, uint32_t sectionPos) {
    if (sections.find((int)sectionId) == sections.end())
        throw std::range_error("Section does not exist: " + std::to_string(sectionId));
    if (sectionPos >= sections[(int)sectionId].size())
        throw std::range_error("Section pos too big. There are " + std::to_string(sections[(int)sectionId].size()) +
                               " and it's trying to access section: " + std::to_string(sectionPos));
    return sections[(int)sectionId][sectionPos].start;
}

uint64_t BinFile::getSectionSize(uint32_t sectionId, uint32_t sectionPos) {
    if (sections.find((int)sectionId) == sections.end())
        throw std::range_error("Section does not exist: " + std::to_string(sectionId));
    if (sectionPos >= sections[(int)sectionId].size())
        throw std::range_error("Section pos too big. There are " + std::to_string(sections[(int)sectionId].size()) +
                               " and it's trying to access section: " + std::to_string(sectionPos));
    return sections[(int)sectionId][sectionPos].size;
}

uint32_t BinFile::readU32LE() {
    if (pos + 4 > size) throw std::range_error("Invalid section size");
    uint32_t r;
    memcpy(&r, (uint8_t *)addr + pos, 4);
    pos += 4;
    return r;
}

uint64_t BinFile::readU64LE() {
    if (pos + 8 > size) throw std::range_error("Invalid section size");
    uint64_t r;
    memcpy(&r, (uint8_t *)addr + pos, 8);
    pos += 8;
    return r;
}

void *BinFile::read(uint64_t len) {
    if (len > size - pos) throw std::range_error("Invalid section size");
    void *r = (uint8_t *)addr + pos;
    pos += len;
    return r;
}

std::unique_ptr<BinFile> openExisting(std::string filename, std::string type, uint32_t maxVersion) {
    return std::unique_ptr<BinFile>(new BinFile(filename, type, maxVersion));
}

}  // namespace BinFileUtils
