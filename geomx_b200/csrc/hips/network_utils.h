// Interface / IP / port discovery (parity: ps-lite src/network_utils.h:28,121,226).
#pragma once
#include <arpa/inet.h>
#include <ifaddrs.h>
#include <net/if.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <cstring>
#include <string>

namespace hips {

inline void GetIP(const std::string& interface, std::string* ip) {
  struct ifaddrs* ifas = nullptr;
  getifaddrs(&ifas);
  for (struct ifaddrs* ifa = ifas; ifa; ifa = ifa->ifa_next) {
    if (!ifa->ifa_addr || ifa->ifa_addr->sa_family != AF_INET) continue;
    if (interface != ifa->ifa_name) continue;
    char buf[INET_ADDRSTRLEN];
    inet_ntop(AF_INET, &reinterpret_cast<struct sockaddr_in*>(ifa->ifa_addr)->sin_addr, buf, INET_ADDRSTRLEN);
    *ip = buf;
    break;
  }
  if (ifas) freeifaddrs(ifas);
}

inline void GetAvailableInterfaceAndIP(std::string* interface, std::string* ip) {
  struct ifaddrs* ifas = nullptr;
  interface->clear(); ip->clear();
  getifaddrs(&ifas);
  for (struct ifaddrs* ifa = ifas; ifa; ifa = ifa->ifa_next) {
    if (!ifa->ifa_addr || ifa->ifa_addr->sa_family != AF_INET) continue;
    if (ifa->ifa_flags & IFF_LOOPBACK) continue;
    char buf[INET_ADDRSTRLEN];
    inet_ntop(AF_INET, &reinterpret_cast<struct sockaddr_in*>(ifa->ifa_addr)->sin_addr, buf, INET_ADDRSTRLEN);
    *ip = buf; *interface = ifa->ifa_name;
    break;
  }
  if (ifas) freeifaddrs(ifas);
}

inline int GetAvailablePort() {
  struct sockaddr_in addr;
  addr.sin_port = htons(0);
  addr.sin_family = AF_INET;
  addr.sin_addr.s_addr = htonl(INADDR_ANY);
  int sock = socket(AF_INET, SOCK_STREAM, 0);
  if (sock < 0) return 0;
  if (bind(sock, reinterpret_cast<struct sockaddr*>(&addr), sizeof(addr)) != 0) { close(sock); return 0; }
  socklen_t len = sizeof(addr);
  if (getsockname(sock, reinterpret_cast<struct sockaddr*>(&addr), &len) != 0) { close(sock); return 0; }
  int port = ntohs(addr.sin_port);
  close(sock);
  return port;
}

}  // namespace hips
