"""soap3dp_b200 -- B200-native GPU alignment hot path of SOAP3-dp."""
